"""Step schedules of the decode engine (mixed into DropoutDecodeEngine, engine.py).

The reference's order per token is two dependent passes over the weights and the K/V cache: one unmasked forward for
y_init, then K masked forwards (models/llava.py:294-305, 336-361).  The schedules here re-order that work without
changing a bit of the result (tests/test_engine_gpu.py compares each with the two-pass engine):

  * speculative single pass   -- K+1 contexts per pass (row 0 unmasked, rows 1..K with the draws as if the relevant set
                                 restored nothing); the K members are repeated only when y_init restores a drawn token;
  * lookahead                 -- that repeat also carries the next step's K+1 contexts, built on the guessed vote;
  * asynchronous batch forms  -- per-sequence step counters, no host decision, one CUDA graph.

State lives on the engine; `_init_schedules` allocates it."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops


class SpeculativeSchedules:
    def _init_schedules(self, speculative: Optional[str], est: int) -> None:
        cfg, B, dev = self.cfg, self.B, self.dev
        vote_on, keep_relevant, aggregate = self.vote_on, self.keep_relevant, self.aggregate
        i32 = dict(dtype=torch.int32, device=dev)
        # ---- speculative single pass (see _spec_head): "auto" = on whenever the K+1 contexts of all sequences fit the
        #      small-batch kernels; DD_SPEC=off|on|auto overrides
        self.speculative = speculative or os.environ.get("DD_SPEC", "auto")
        if self.speculative not in ("auto", "on", "off"):
            raise ValueError(self.speculative)
        ok = self.K + 1 <= 16        # (more than 32 query rows per kv head: the attention runs in member groups)
        # one sequence: the host decides per step whether to fall back (the hidden-state vote of InstructBLIP always
        # takes the asynchronous schedule below, also at B = 1)
        fits = ok and B == 1 and vote_on == "logits"
        if fits and keep_relevant and 16 < 2 * self.K + 1 <= 32 and os.environ.get("DD_BATCH_LA", "auto") != "0":
            fits = False    # (K = 8: the 17 slots of the lookahead exceed the <= 16-row kernels; the batch schedule has it)
        if self.speculative == "on" and not ok:
            raise ops.DropDecError("speculative single pass needs K + 1 <= 16 contexts")
        self._spec_fits = fits
        self.spec_active = fits and self.speculative != "off"
        # batches: some sequence nearly always needs the fallback, so sequences advance independently instead
        # (asynchronous batch schedule, _spec_batch_step; per-sequence flags on the device, no host decision)
        self.spec_batch = ok and not fits and self.speculative != "off"
        self.n_spec_steps = self.n_fallbacks = 0        # cumulative over the engine's life
        self._seq_steps_done = 0
        self._spec_base = (0, 0)                        # counters when the current sequence batch started
        self._last_src = "pb"
        self._graphs: Optional[dict] = None
        if self.spec_batch:
            V = cfg.vocab
            self.seq_steps = torch.zeros(B, **i32)          # ensemble steps done, per sequence
            self.in_fb = torch.zeros(B, **i32)              # 1: the next pass repeats this step with the real masks
            self.target = torch.zeros(1, **i32)
            self.active = torch.zeros(B, **i32)
            self.adv = torch.zeros(B, **i32)
            self._fb_total = torch.zeros(1, dtype=torch.int64, device=dev)
            self._mask_tmp = torch.zeros(B, self.K, self.words, **i32)
            self._mask_real = torch.zeros(B, self.K, self.words, **i32)
            self._nm_real = torch.zeros(B, self.K, **i32)
            self._ids_c = torch.zeros(B, self.K, **i32)
            self._win_c, self._tok_sc = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._minus1 = torch.full((B,), -1, **i32)
            self._zero = torch.zeros(B, **i32)
            self.ret_logits = torch.zeros(B, V, dtype=torch.float32, device=dev)
            if aggregate == "mean":
                self._out_c = torch.zeros(B, V, dtype=torch.float32, device=dev)
            if vote_on == "hidden":
                self._hid32c = torch.zeros(B, self.K + 1, cfg.d_model, dtype=torch.float32, device=dev)
        # batches with lookahead (_spec_batch_la_step): 2K+1 context slots per sequence.  Measured on the bench workload
        # (7B, fallback rate 0.37-0.41, tokens/s with / without): B=3 591/446, B=4 742/555, B=8 1502/1158, B=16 2435/2097,
        # B=32 3226/3138; LLaVA-NeXT K=8 B=2 271/203; but B=2 K=3 399/426 (14 rows leave the fused <=16-row kernels) and
        # no gain once the pass is compute-bound -> auto = 16 < rows <= 256.  DD_BATCH_LA=0|1 overrides.
        bla = os.environ.get("DD_BATCH_LA", "auto")
        self.batch_la = bool(self.spec_batch and vote_on == "logits" and keep_relevant and 2 * self.K + 1 <= 32
                             and (bla == "1" or (bla == "auto" and 16 < B * (2 * self.K + 1) <= 256)))
        if self.batch_la:
            K, M = self.K, 2 * self.K + 1
            self.pe = self._alloc_pass(M, est)
            self.la_mask = torch.zeros(B, M, self.words, **i32)
            self._g_tok, self._g_win = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._gw = torch.zeros(B, **i32)
            self._steps_spec, self._seq_p = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._row_tok, self._row_pos = torch.zeros(B * M, **i32), torch.zeros(B * M, **i32)
            self._act1, self._act2 = torch.zeros(B, **i32), torch.zeros(B, **i32)
            # both member windows in ONE attention launch (K/V read once for all 2K+1 contexts) when the pass is
            # unsplit and the contexts fit the kernel's 32 query rows per kv head; DD_LA_ONE_ATTN=0 keeps two launches
            self._len_extra = torch.zeros(B, M, **i32)
            self._la_one_attn = (self.pe.splits == 1 and M * (cfg.n_heads // cfg.n_kv_heads) <= 32
                                 and os.environ.get("DD_LA_ONE_ATTN", "1") == "1")
            self._y_new = torch.zeros(B, **i32)
            self._ids_v, self._ids_s = torch.zeros(B, K, **i32), torch.zeros(B, K + 1, **i32)
            self._win_v, self._tok_v = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._win_s, self._tok_s = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._w1, self._w2, self._tok1 = torch.zeros(B, **i32), torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._adv2 = torch.zeros(B, **i32)
            if aggregate == "mean":
                self._out_v = torch.zeros(B, cfg.vocab, dtype=torch.float32, device=dev)
                self._out1 = torch.zeros(B, cfg.vocab, dtype=torch.float32, device=dev)
        if self.spec_active or self.spec_batch:
            self.pc = self._alloc_pass(self.K + 1, est)
            self.spec_draw = torch.zeros(B, self.K, self.words, **i32)
            self.spec_mask = torch.zeros(B, self.K + 1, self.words, **i32)       # member 0 = the unmasked context
            self._nm_spec = torch.zeros(B, self.K, **i32)
            self.member_ids_c = torch.zeros(B, self.K + 1, **i32)
            self.winner_c = torch.zeros(B, **i32)
            self._tok_c = torch.zeros(B, **i32)
            self.winner_p1 = torch.zeros(B, **i32)
            self._spec_flags = torch.zeros(2, **i32)                # dd_schedule_flags: {-, the step must be repeated}
            self.need_fb = self._spec_flags[1:2]
            self.need_fb_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            if self.spec_active:
                # one sequence: ONE mask buffer [real masks of the step (K) | unmasked context | draws (K)] -- the
                # speculative pass reads rows K..2K, the lookahead pass all 2K+1, nothing is copied between them
                self._mask_all = torch.zeros(B, 2 * self.K + 1, self.words, **i32)
                self.bitmask = self._mask_all[:, :self.K]
                self.spec_mask = self._mask_all[:, self.K:]
                self.spec_draw = self._mask_all[:, self.K + 1:]
            self.tok_host = torch.zeros(B, dtype=torch.int32).pin_memory()     # input token of the pass (EOS check)
            self.tok2_host = torch.zeros(B, dtype=torch.int32).pin_memory()    # token a lookahead pass committed (EOS check)
        # ---- lookahead inside the fallback pass (one sequence, see _spec_chain): 2K+1 contexts
        la = os.environ.get("DD_LOOKAHEAD", "1") == "1"
        self.lookahead = bool(la and fits and self.speculative != "off" and keep_relevant and 2 * self.K + 1 <= 16)
        self.n_chain = self.n_chain_ok = 0
        if self.lookahead:
            K = self.K
            self.pd = self._alloc_pass(2 * K + 1, est)
            self.chain_mask = self._mask_all
            self._g_tok, self._g_win = torch.zeros(B, **i32), torch.zeros(B, **i32)     # the guessed outcome of step t
            self._g_ids = torch.zeros(B, K, **i32)
            self._seq_p1, self._step_p1 = torch.zeros(B, **i32), torch.zeros(1, **i32)
            self._row_tok_c, self._row_pos_c = torch.zeros(2 * K + 1, **i32), torch.zeros(2 * K + 1, **i32)
            self._ids_d = torch.zeros(B, K + 1, **i32)
            self._win_d, self._tok_d = torch.zeros(B, **i32), torch.zeros(B, **i32)
            self._win_pk = torch.zeros(B, **i32)
            self.chain_flags = torch.zeros(2, **i32)                # {guess was right, step t+1 needs its fallback}
            self.chain_flags_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._g_out = (torch.zeros(B, cfg.vocab, dtype=torch.float32, device=dev) if aggregate == "mean" else None)

    # ------------------------------------------------------------------------------------
    # Speculative single pass.  The reference needs the preliminary forward only for y_init = argmax(unmasked
    # logits), and y_init enters the masks only through the relevant set S = {i : y_init in topk_ids[i]} whose
    # tokens are un-dropped (models/llava.py:443-482, 657-660).  The draws themselves do not depend on y_init.  So
    # one pass runs K+1 contexts at once -- row 0 unmasked, rows 1..K with the draws as if S restored nothing -- at
    # the cost of one weight stream instead of two.  y_init then comes out of row 0, the real masks are sampled, and
    # if they equal the speculated ones (S empty, or none of its tokens was drawn: the common case) rows 1..K ARE
    # the reference's K masked forwards, bit for bit.  Otherwise the masked pass is re-run with the real masks
    # (the reference's own order).  With one sequence the host takes that decision per token (replacing the
    # reference's >= 3K syncs); batches use the asynchronous schedule of _spec_batch_step instead.
    # ------------------------------------------------------------------------------------
    def _spec_head(self) -> None:
        self.launches_per_step = 0
        u = self._step_uniforms()
        self._sample(u, False, self.spec_draw, self._nm_spec)        # (the draws ARE rows 1..K of spec_mask)
        logits_c = self._forward(self.pc, self.spec_mask)
        ops.aggregate(logits_c, "vote", member_ids=self.member_ids_c, winner=self.winner_c, token=self._tok_c,
                      vocab=self.cfg.vocab)                                        # per-row argmax of all K+1 contexts
        self.launches_per_step += 1
        self.y_init.copy_(self.member_ids_c[:, 0])
        if self.keep_relevant:
            self._sample(u, True, self.bitmask, self.n_masked)                     # the reference's masks
            ops.schedule_flags(self._spec_flags, mask_a=self.bitmask, mask_b=self.spec_draw)   # need_fb
        else:                                                                      # epis_no_overlap: nothing to restore
            self.bitmask.copy_(self.spec_draw)
            self.n_masked.copy_(self._nm_spec)
            self.need_fb.zero_()
        if self.lookahead:
            self._guess_from(self.pc.logits, 1)
        self.tok_host.copy_(self.cur_token, non_blocking=True)
        self.need_fb_host.copy_(self.need_fb, non_blocking=True)

    # ------------------------------------------------------------------------------------
    # Lookahead inside the fallback pass (one sequence).  When y_init restored a drawn token, step t needs a second
    # weight stream for its K real-mask members anyway.  That stream also carries the K+1 speculative contexts of step
    # t+1, built on a GUESS of step t's outcome: the token and the winning member voted by the wrong-mask members of
    # the pass before (they differ from the real members only in the few restored tokens).  Rows 0..K-1: the members of
    # step t (real masks, position S).  After the QKV projection of a layer the guessed winner's new K/V row is written
    # to cache row S (the real commit overwrites it later); rows K..2K -- unmasked + the draws of step t+1, input =
    # the guessed token, position S+1 -- then attend to rows [0, S] exactly as they would after a commit.  If the
    # real vote confirms the guess, rows K..2K ARE step t+1's speculative pass, bit for bit, and its weight stream has
    # been saved; if not, they are dropped.  Only views and existing kernels: per-row token / RoPE position tensors (the
    # kernels see 2K+1 one-member sequences) and two attention windows.
    # ------------------------------------------------------------------------------------
    def _guess_from(self, logits: torch.Tensor, first: int) -> None:
        """Guess step t's (token, winner) from the wrong-mask members = rows first..first+K-1 of `logits`."""
        ops.aggregate(logits, self.aggregate, member_ids=self._g_ids, winner=self._g_win, token=self._g_tok,
                      out_logits=self._g_out, vocab=self.cfg.vocab, first_member=first, members=self.K)
        self.launches_per_step += 1

    def _attention_rows(self, pb, l: int, bitmask, seq_lens, m0: int, m1: int, seq_active=None) -> None:
        G = self.cfg.n_heads // self.cfg.n_kv_heads
        per = max(1, 32 // G)
        for a in range(m0, m1, per):
            ops.decode_attention(pb.q, pb.k_new[l], pb.v_new[l], self.k_cache[l], self.v_cache[l], seq_lens,
                                 self.vis_start, bitmask, self.n_vis, self.scale, pb.splits, pb.attn, pb.ws,
                                 seq_active=seq_active, member_offset=a, members=min(per, m1 - a))
            self.launches_per_step += 1 + (pb.splits == 1)

    def _chain_forward_launches(self, pd, K: int, M: int) -> None:
        """The forward of the chain pass as a chain of launches (passes dd_decoder_pass does not serve)."""
        cfg = self.cfg
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        ops.embed_gather(self.w.embed, self._row_tok_c, 1, pd.x)
        x, x2 = pd.x, pd.x2
        G = H // Hkv
        two_windows = (pd.splits >= 2 and (K + 1) * G <= 32 and os.environ.get("DD_TWO_WINDOW_ATTN", "1") == "1")
        if two_windows and pd.ws2 is None:
            pd.ws2 = ops.two_window_attention_workspace(self.B, M, H, D, pd.splits, self.dev)
        for l, lw in enumerate(self.w.layers):
            ops.rmsnorm(x, lw["ln1"], cfg.rms_eps, pd.xn)
            # (the guessed winner's new K/V row goes to cache row S from the projection's epilogue: no commit launch)
            ops.decoder_linear(pd.xn, lw["wqkv"], ops.EPI_ROPE,
                               rope=dict(cos=self.cos, sin=self.sin, seq_lens=self._row_pos_c, members=1, H=H, H_kv=Hkv,
                                         D=D, q_out=pd.q, k_new=pd.k_new[l], v_new=pd.v_new[l],
                                         guess_winner=self._g_win, guess_pos=self.seq_lens, guess_members=M,
                                         k_cache=self.k_cache[l], v_cache=self.v_cache[l]))
            if two_windows:
                # both member windows in ONE launch (split key ranges, each with its own partials): side by side instead
                # of one latency-bound launch after the other
                ops.decode_attention_two_windows(pd.q, pd.k_new[l], pd.v_new[l], self.k_cache[l], self.v_cache[l],
                                                 self.seq_lens, self._seq_p1, K, self.vis_start, self.chain_mask,
                                                 self.n_vis, self.scale, pd.splits, pd.attn, pd.ws2)
                self.launches_per_step += 1
            else:
                self._attention_rows(pd, l, self.chain_mask, self.seq_lens, 0, K)
                self._attention_rows(pd, l, self.chain_mask, self._seq_p1, K, M)
            ops.decoder_linear(pd.attn.view(M, H * D), lw["wo"], ops.EPI_RESIDUAL, residual=x, out=x2)
            ops.rmsnorm(x2, lw["ln2"], cfg.rms_eps, pd.xn)
            ops.decoder_linear(pd.xn, lw["wgu"], ops.EPI_SWIGLU, out=pd.act)
            ops.decoder_linear(pd.act, lw["wd"], ops.EPI_RESIDUAL, residual=x2, out=x)
            self.launches_per_step += 6
        ops.rmsnorm(x, self.w.norm, cfg.rms_eps, pd.xn)
        ops.decoder_linear(pd.xn, self.w.lm_head_padded, ops.EPI_F32, out=pd.logits.view(M, self.Vp))

    def _spec_chain(self) -> None:
        self.launches_per_step = 0
        cfg, K, pd = self.cfg, self.K, self.pd
        H, Hkv, D, M = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, 2 * self.K + 1
        self.tok_host.copy_(self.cur_token, non_blocking=True)
        torch.add(self.seq_lens, 1, out=self._seq_p1)
        torch.add(self.step_ctr, 1, out=self._step_p1)
        self._sample(self._step_uniforms(1), False, self.spec_draw, self._nm_spec, step=self._step_p1)   # draws of t+1
        # (chain_mask = [bitmask | unmasked | spec_draw] is one buffer: nothing to copy)
        # per-row input token and RoPE position: the kernels see 2K+1 "sequences" of one member each
        self._row_tok_c[:K] = self.cur_token
        self._row_tok_c[K:] = self._g_tok
        self._row_pos_c[:K] = self.seq_lens
        self._row_pos_c[K:] = self._seq_p1
        self._set_chain(M)
        if self._pass_kernel_ok(pd, M, K + 1):
            # one persistent kernel: per-row tokens / positions, two member windows, the guessed winner's row written
            # to cache row S by the QKV epilogue
            self._run_pass_kernel(pd, tok=self._row_tok_c, tok_div=1, pos=self._row_pos_c, pos_div=1,
                                  bitmask=self.chain_mask, windows=[(0, K, self.seq_lens), (K, K + 1, self._seq_p1)],
                                  guess_winner=self._g_win, guess_pos=self.seq_lens)
        else:
            self._chain_forward_launches(pd, K, M)
        V, mean = cfg.vocab, self.aggregate == "mean"
        # ---- step t: the reference's vote over its K real-mask members (rows 0..K-1)
        ops.aggregate(pd.logits, self.aggregate, member_ids=self.member_ids, winner=self.winner, token=self.next_token,
                      out_logits=self.out_logits if mean else None, vocab=V, first_member=0, members=K)
        if mean:
            ops.token_stats(self.out_logits, V, self.token_stats, slot=self.step_ctr, slot_offset=1)
        else:
            ops.token_stats(pd.logits, V, self.token_stats, member=self.winner, slot=self.step_ctr, slot_offset=1)
        ops.commit_winner(pd.k_new, pd.v_new, self.winner, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
        self.cur_token.copy_(self.next_token)
        self.tok2_host.copy_(self.next_token, non_blocking=True)   # step t's result: the host must not chain past an EOS
        # ---- step t+1 as _spec_head would see it (only meaningful when the guess was right)
        ops.aggregate(pd.logits, "vote", member_ids=self._ids_d, winner=self._win_d, token=self._tok_d, vocab=V,
                      first_member=K, members=K + 1)
        self.y_init.copy_(self._ids_d[:, 0])
        self._sample(self._step_uniforms(0), True, self.bitmask, self.n_masked)            # the reference's masks of t+1
        # {the guessed (token, winner) of step t was right, step t+1 must be repeated}: one launch
        ops.schedule_flags(self.chain_flags, mask_a=self.bitmask, mask_b=self.spec_draw, tok_a=self.next_token,
                           tok_b=self._g_tok, win_a=self.winner, win_b=self._g_win)
        self.launches_per_step += 11
        self._guess_from(pd.logits, K + 1)          # should step t+1 fall back too: its own guess
        self.chain_flags_host.copy_(self.chain_flags, non_blocking=True)

    def _spec_fin_chain(self) -> None:
        """Step t+1 completed by the lookahead rows: members = rows K+1..2K of the chain pass."""
        self.launches_per_step = 0
        K, pd = self.K, self.pd
        ops.aggregate(pd.logits, self.aggregate, member_ids=self.member_ids, winner=self.winner,
                      token=self.next_token, out_logits=self.out_logits if self.aggregate == "mean" else None,
                      vocab=self.cfg.vocab, first_member=K + 1, members=K)
        torch.add(self.winner, K + 1, out=self._win_pk)
        if self.aggregate == "mean":
            ops.token_stats(self.out_logits, self.cfg.vocab, self.token_stats, slot=self.step_ctr, slot_offset=1)
        else:
            ops.token_stats(pd.logits, self.cfg.vocab, self.token_stats, member=self._win_pk, slot=self.step_ctr,
                            slot_offset=1)
        ops.commit_winner(pd.k_new, pd.v_new, self._win_pk, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
        self.cur_token.copy_(self.next_token)
        self.launches_per_step += 5

    def _spec_fin(self) -> None:
        """The speculated masks were the real ones: rows 1..K of the speculative pass are the K members."""
        self.launches_per_step = 0
        ops.aggregate(self.pc.logits, self.aggregate, member_ids=self.member_ids, winner=self.winner,
                      token=self.next_token, out_logits=self.out_logits if self.aggregate == "mean" else None,
                      vocab=self.cfg.vocab, first_member=1, members=self.K)
        torch.add(self.winner, 1, out=self.winner_p1)
        if self.aggregate == "mean":
            ops.token_stats(self.out_logits, self.cfg.vocab, self.token_stats, slot=self.step_ctr, slot_offset=1)
        else:
            ops.token_stats(self.pc.logits, self.cfg.vocab, self.token_stats, member=self.winner_p1, slot=self.step_ctr,
                            slot_offset=1)
        ops.commit_winner(self.pc.k_new, self.pc.v_new, self.winner_p1, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance(self.next_token, self.tokens_out, self.seq_lens, self.step_ctr)
        self.cur_token.copy_(self.next_token)
        self.launches_per_step += 4

    def _spec_fallback(self) -> None:
        self.launches_per_step = 0
        self._masked_pass_and_commit()

    def _spec_batch_step(self) -> None:
        """Asynchronous batch schedule (B >= 2; no host decision, ONE pass per call): every sequence runs K+1
        contexts -- row 0 unmasked, rows 1..K with the draws of ITS current step, or, if its previous pass found
        that y_init restores a drawn token, with the real masks of that step.  A sequence whose rows 1..K carried
        its real masks votes, commits and advances; the others repeat the step in the next pass.  Sequences are
        independent (per-sequence step counters key the Philox stream and every per-step buffer), so each one
        produces exactly the tokens of the two-pass order while the batch pays ~(1 + fallback rate) weight and
        K/V streams per token instead of two."""
        self.launches_per_step = 0
        V, K, mean = self.cfg.vocab, self.K, self.aggregate == "mean"
        steps = self.seq_steps
        u = None
        if self.uniforms is not None:       # parity mode: sequence b uses the pre-drawn numbers of ITS step
            idx = steps.long().clamp(max=self.uniforms.shape[0] - 1)
            self._u_step.copy_(self.uniforms[idx, self._rows()])
            u = self._u_step
        self._sample(u, False, self.spec_draw, self._nm_spec, per_seq=True)
        infb = self.in_fb.bool()
        self._mask_tmp.copy_(torch.where(infb[:, None, None], self.bitmask, self.spec_draw))
        self.spec_mask[:, 1:].copy_(self._mask_tmp)
        act = steps < self.target
        self.active.copy_(act.to(torch.int32))
        logits_c = self._forward(self.pc, self.spec_mask, seq_active=self.active)
        ops.aggregate(logits_c, "vote", member_ids=self.member_ids_c, winner=self.winner_c, token=self._tok_c, vocab=V)
        # y_init / bitmask / n_masked stay what they were for sequences that are waiting at the target (they keep
        # describing the sequence's last completed step) and for y_init of a sequence that is repeating its step
        self.y_init.copy_(torch.where(act & ~infb, self.member_ids_c[:, 0], self.y_init))
        if self.keep_relevant:
            self._sample(u, True, self._mask_real, self._nm_real, per_seq=True)            # the reference's masks
            ok = infb | (self._mask_real == self._mask_tmp).flatten(1).all(1)
        else:                                                                          # epis_no_overlap
            self._mask_real.copy_(self.spec_draw); self._nm_real.copy_(self._nm_spec)
            ok = torch.ones_like(infb)
        self.bitmask.copy_(torch.where(act[:, None, None], self._mask_real, self.bitmask))
        self.n_masked.copy_(torch.where(act[:, None], self._nm_real, self.n_masked))
        adv = ok & act
        self.adv.copy_(adv.to(torch.int32))
        if self.vote_on == "hidden":
            # InstructBLIP (models/instructblip.py:125-140): the vote is taken on argmax of the members' final hidden
            # states; the returned logits / token are lm_head of the winner = its row of the pass's logits
            self._hid32c.copy_(self.pc.xn.view(self.B, K + 1, self.cfg.d_model))
            ops.aggregate(self._hid32c, "vote", member_ids=self._ids_c, winner=self._win_c, token=self._tok_sc,
                          first_member=1, members=K)
            self._tok_sc.copy_(self.member_ids_c.gather(1, (self._win_c + 1).long()[:, None])[:, 0])
        else:
            ops.aggregate(self.pc.logits, self.aggregate, member_ids=self._ids_c, winner=self._win_c,
                          token=self._tok_sc, out_logits=self._out_c if mean else None, vocab=V, first_member=1,
                          members=K)
        self.member_ids.copy_(torch.where(adv[:, None], self._ids_c, self.member_ids))
        self.winner.copy_(torch.where(adv, self._win_c, self.winner))
        self.next_token.copy_(torch.where(adv, self._tok_sc, self.next_token))
        self.winner_p1.copy_(torch.where(adv, self._win_c + 1, self._minus1))
        if mean:
            self.out_logits.copy_(torch.where(adv[:, None], self._out_c, self.out_logits))
            ops.token_stats(self._out_c, V, self.token_stats, member=torch.where(adv, self._zero, self._minus1),
                            slot=steps, slot_offset=1, slot_per_seq=True)
            self.ret_logits.copy_(self.out_logits)
        else:
            ops.token_stats(self.pc.logits, V, self.token_stats, member=self.winner_p1, slot=steps, slot_offset=1,
                            slot_per_seq=True)
            rows = self.pc.logits[self._rows(), self.winner_p1.clamp(min=0).long(), :V]
            self.ret_logits.copy_(torch.where(adv[:, None], rows, self.ret_logits))
        ops.commit_winner(self.pc.k_new, self.pc.v_new, self.winner_p1, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance_seq(self._tok_sc, self.adv, self.tokens_out, self.seq_lens, steps)
        self.cur_token.copy_(torch.where(adv, self._tok_sc, self.cur_token))
        self.in_fb.copy_((~ok & act).to(torch.int32))
        self._fb_total.add_(self.in_fb.sum())
        self.launches_per_step += 7

    def _spec_batch_la_step(self) -> None:
        """Asynchronous batch schedule with lookahead: 2K+1 context slots per sequence and pass.  Slots 0..K-1 ("verify")
        run the real-mask members of a step whose speculated masks were wrong (idle otherwise); slots K..2K ("spec")
        run the speculative contexts -- unmasked + draws -- of the sequence's next undecided step: the same step for a
        sequence that is not repeating one, the FOLLOWING step, on the guessed outcome of the verify slots (token and
        winning member voted by the wrong-mask members of the pass before; the guessed winner's K/V row is written
        ahead into the cache, layer by layer), for one that is.  A confirmed guess makes the spec slots that following
        step's speculative pass, so a sequence can complete two steps in one pass and a repeated step costs no extra
        weight stream.  Everything is decided per sequence on the device; results equal the two-pass order bit for bit."""
        self.launches_per_step = 0
        cfg, B, K, pe = self.cfg, self.B, self.K, self.pe
        V, M, mean = cfg.vocab, 2 * self.K + 1, self.aggregate == "mean"
        H, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        steps = self.seq_steps
        act = steps < self.target
        fb = self.in_fb.bool() & act
        fb_i = fb.to(torch.int32)
        self._steps_spec.copy_(steps + fb_i)
        spec_on = act & (self._steps_spec < self.target)
        u = None
        if self.uniforms is not None:       # parity mode: the pre-drawn numbers of the step in the spec slots
            idx = self._steps_spec.long().clamp(max=self.uniforms.shape[0] - 1)
            self._u_step.copy_(self.uniforms[idx, self._rows()])
            u = self._u_step
        self._sample(u, False, self.spec_draw, self._nm_spec, per_seq=True, step=self._steps_spec)
        self.la_mask[:, :K].copy_(self.bitmask)
        self.la_mask[:, K + 1:].copy_(self.spec_draw)
        rt, rp = self._row_tok.view(B, M), self._row_pos.view(B, M)
        rt[:, :K] = self.cur_token[:, None]
        rt[:, K:] = torch.where(fb, self._g_tok, self.cur_token)[:, None]
        self._seq_p.copy_(self.seq_lens + fb_i)
        rp[:, :K] = self.seq_lens[:, None]
        rp[:, K:] = self._seq_p[:, None]
        self._act1.copy_(fb_i)
        self._act2.copy_(spec_on.to(torch.int32))
        self._gw.copy_(torch.where(fb, self._g_win, self._minus1))
        if self._la_one_attn:
            self._len_extra[:, K:] = fb_i[:, None]      # the spec slots of a repeating sequence see the row committed ahead
        # ---- the pass
        rows = B * M
        self._set_chain(rows)
        ops.embed_gather(self.w.embed, self._row_tok, 1, pe.x)
        x, x2 = pe.x, pe.x2
        for l, lw in enumerate(self.w.layers):
            ops.rmsnorm(x, lw["ln1"], cfg.rms_eps, pe.xn)
            self._linear(pe.xn, lw["wqkv"], ops.EPI_NONE, pe.qkv)
            ops.rope_stage(pe.qkv, self.cos, self.sin, self._row_pos, rows, 1, H, Hkv, D, pe.q, pe.k_new[l], pe.v_new[l])
            ops.commit_winner(pe.k_new[l:l + 1], pe.v_new[l:l + 1], self._gw, self.seq_lens, self.k_cache[l:l + 1],
                              self.v_cache[l:l + 1])                                   # the guessed winner's row, ahead
            if self._la_one_attn:
                ops.decode_attention(pe.q, pe.k_new[l], pe.v_new[l], self.k_cache[l], self.v_cache[l], self.seq_lens,
                                     self.vis_start, self.la_mask, self.n_vis, self.scale, 1, pe.attn, pe.ws,
                                     seq_active=self._act1, seq_active2=self._act2, len_extra=self._len_extra)
                self.launches_per_step += 2
            else:
                self._attention_rows(pe, l, self.la_mask, self.seq_lens, 0, K, seq_active=self._act1)
                self._attention_rows(pe, l, self.la_mask, self._seq_p, K, M, seq_active=self._act2)
            if rows <= 32 and self.gemm_backend == "auto":
                ops.skinny_gemm(pe.attn.view(rows, H * D), lw["wo"], ops.EPI_RESIDUAL, residual=x, out=x2)
                self.launches_per_step += 1
            else:
                self._linear(pe.attn.view(rows, H * D), lw["wo"], ops.EPI_RESIDUAL, x2, residual=x)
            ops.rmsnorm(x2, lw["ln2"], cfg.rms_eps, pe.xn)
            self._linear(pe.xn, lw["wgu"], ops.EPI_SWIGLU, pe.act)
            self._linear(pe.act, lw["wd"], ops.EPI_RESIDUAL, x, residual=x2)
            self.launches_per_step += 4
        ops.rmsnorm(x, self.w.norm, cfg.rms_eps, pe.xn)
        self._linear(pe.xn, self.w.lm_head_padded, ops.EPI_F32, pe.logits.view(rows, self.Vp))
        self.launches_per_step += 2
        # ---- votes: verify slots, per-row argmax of the spec slots, members among the spec slots
        ops.aggregate(pe.logits, self.aggregate, member_ids=self._ids_v, winner=self._win_v, token=self._tok_v,
                      out_logits=self._out_v if mean else None, vocab=V, first_member=0, members=K)
        ops.aggregate(pe.logits, "vote", member_ids=self._ids_s, winner=self._win_s, token=self._tok_s, vocab=V,
                      first_member=K, members=K + 1)
        self._y_new.copy_(self._ids_s[:, 0])
        self._sample(u, True, self._mask_real, self._nm_real, per_seq=True, step=self._steps_spec, y=self._y_new)
        eq = (self._mask_real == self.spec_draw).flatten(1).all(1)
        ops.aggregate(pe.logits, self.aggregate, member_ids=self._ids_c, winner=self._win_c, token=self._tok_sc,
                      out_logits=self._out_c if mean else None, vocab=V, first_member=K + 1, members=K)
        self.launches_per_step += 5
        chain_ok = fb & spec_on & (self._tok_v == self._g_tok) & (self._win_v == self._g_win)
        valid = (act & ~fb) | chain_ok              # the spec slots were a real speculative pass of step _steps_spec
        adv1 = fb | (valid & ~fb & eq)
        adv2 = chain_ok & eq
        win_m = self._win_c + (K + 1)
        self._w1.copy_(torch.where(adv1, torch.where(fb, self._win_v, win_m), self._minus1))
        self._w2.copy_(torch.where(adv2, win_m, self._minus1))
        self._tok1.copy_(torch.where(fb, self._tok_v, self._tok_sc))
        self.adv.copy_(adv1.to(torch.int32))
        self._adv2.copy_(adv2.to(torch.int32))
        # ---- per-step views: the step in the spec slots (completed or to be repeated) / the last completed step
        self.y_init.copy_(torch.where(valid, self._y_new, self.y_init))
        self.bitmask.copy_(torch.where(valid[:, None, None], self._mask_real, self.bitmask))
        self.n_masked.copy_(torch.where(valid[:, None], self._nm_real, self.n_masked))
        ids1 = torch.where(fb[:, None], self._ids_v, self._ids_c)
        self.member_ids.copy_(torch.where(adv2[:, None], self._ids_c, torch.where(adv1[:, None], ids1, self.member_ids)))
        self.winner.copy_(torch.where(adv2, self._win_c, torch.where(adv1, torch.where(fb, self._win_v, self._win_c),
                                                                      self.winner)))
        self.next_token.copy_(torch.where(adv2, self._tok_sc, torch.where(adv1, self._tok1, self.next_token)))
        self.winner_p1.copy_(torch.where(adv2, self._w2, self._w1))       # row of the pass that holds the returned logits
        if mean:
            self._out1.copy_(torch.where(fb[:, None], self._out_v, self._out_c))
            self.out_logits.copy_(torch.where(adv2[:, None], self._out_c,
                                              torch.where(adv1[:, None], self._out1, self.out_logits)))
            self.ret_logits.copy_(self.out_logits)
            ops.token_stats(self._out1, V, self.token_stats, member=torch.where(adv1, self._zero, self._minus1),
                            slot=steps, slot_offset=1, slot_per_seq=True)
        else:
            r = self.pe.logits[self._rows(), self.winner_p1.clamp(min=0).long(), :V]
            self.ret_logits.copy_(torch.where((adv1 | adv2)[:, None], r, self.ret_logits))
            ops.token_stats(pe.logits, V, self.token_stats, member=self._w1, slot=steps, slot_offset=1,
                            slot_per_seq=True)
        ops.commit_winner(pe.k_new, pe.v_new, self._w1, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance_seq(self._tok1, self.adv, self.tokens_out, self.seq_lens, steps)
        if mean:
            ops.token_stats(self._out_c, V, self.token_stats, member=torch.where(adv2, self._zero, self._minus1),
                            slot=steps, slot_offset=1, slot_per_seq=True)
        else:
            ops.token_stats(pe.logits, V, self.token_stats, member=self._w2, slot=steps, slot_offset=1,
                            slot_per_seq=True)
        ops.commit_winner(pe.k_new, pe.v_new, self._w2, self.seq_lens, self.k_cache, self.v_cache)
        ops.step_advance_seq(self._tok_sc, self._adv2, self.tokens_out, self.seq_lens, steps)
        self.cur_token.copy_(torch.where(adv2, self._tok_sc, torch.where(adv1, self._tok1, self.cur_token)))
        self._g_tok.copy_(self._tok_sc)
        self._g_win.copy_(self._win_c)
        self.in_fb.copy_((valid & ~eq).to(torch.int32))
        self._fb_total.add_(self.in_fb.sum())
        self.launches_per_step += 6

    def fallback_rate(self) -> float:
        """Fraction of (step, sequence) pairs whose masked pass had to be re-run with the real masks."""
        if self.spec_batch:         # repeated (step, sequence) pairs / completed ones
            return float(self._fb_total.item()) / max(1, self._seq_steps_done)
        return self.n_fallbacks / max(1, self.n_spec_steps)

    def _run_piece(self, graphs: Optional[dict], name: str, fn) -> None:
        if graphs is not None:
            graphs[name][0].replay()
            self.launch_count += graphs[name][1]
        else:
            fn()
            self.launch_count += self.launches_per_step

    def _spec_steps(self, n_steps: int, graphs: Optional[dict], stop_token: Optional[int] = None) -> int:
        """n_steps ensemble steps through the speculative path (one sequence); `graphs` = captured pieces or None
        (eager).  One host decision (a stream synchronisation) per pass.  Lookahead state never outlives the call.
        stop_token: every pass also hands the host the token it started from, so the loop ends -- at no extra
        synchronisation, at most one pass late -- once that token has been generated (EOS).  Returns the steps done."""
        run = lambda name, fn: self._run_piece(graphs, name, fn)
        sync = torch.cuda.current_stream().synchronize
        i, src, need_fb = 0, None, False        # src: which pass already served as step i's speculative pass
        while i < n_steps:
            if not self.spec_active:            # the adaptive switch fired: the reference's own order
                if stop_token is not None:      # (one synchronisation per token in this mode, like every other pass)
                    self.tok_host.copy_(self.cur_token, non_blocking=True)
                    sync()
                    if int(self.tok_host[0]) == stop_token:
                        return i
                if graphs is not None:
                    if self._graph is None:
                        self.capture()
                    self._graph.replay()
                else:
                    self._two_pass_step()
                self.launch_count += self.launches_per_step
                self._last_src = "pb"
                i += 1
                continue
            if src is None:
                run("head", self._spec_head)
                sync()
                if stop_token is not None and int(self.tok_host[0]) == stop_token:
                    return i                    # the previous step produced EOS; this pass committed nothing
                need_fb, src = bool(int(self.need_fb_host[0])), "pc"
            self.n_spec_steps += 1
            if not need_fb:
                if src == "pc":
                    run("fin", self._spec_fin)
                    self._last_src = "pc"
                else:
                    run("fin_chain", self._spec_fin_chain)
                    self._last_src = "pd1"
                src = None
            else:
                self.n_fallbacks += 1
                if self.lookahead and n_steps - i >= 2:
                    run("chain", self._spec_chain)
                    sync()
                    self.n_chain += 1
                    self._last_src = "pd0"
                    if stop_token is not None and stop_token in (int(self.tok_host[0]), int(self.tok2_host[0])):
                        # EOS came out of the step before (this pass has committed one more), or out of the step this
                        # pass completed: either way stop here -- the lookahead rows must not be finished past an EOS
                        return i + 1
                    if int(self.chain_flags_host[0]):
                        self.n_chain_ok += 1
                        need_fb, src = bool(int(self.chain_flags_host[1])), "pd"
                    else:
                        src = None
                else:
                    run("fallback", self._spec_fallback)
                    self._last_src = "pb"
                    src = None
            i += 1
            # adaptive: when almost every step needs the fallback (a checkpoint whose next token is nearly always one
            # of the visual tokens' top-k ids) and the lookahead cannot absorb it, the reference's order is cheaper
            n, f = self.n_spec_steps - self._spec_base[0], self.n_fallbacks - self._spec_base[1]
            if self.speculative == "auto" and not self.lookahead and n >= 32 and f > 0.8 * n and src is None:
                self.spec_active = False
        return i

    def _decode_async(self, n_steps: int, use_graph: bool) -> None:
        """Passes of the asynchronous batch schedule until every sequence has advanced n_steps: n_steps passes, then
        as many more as the sequence with the most repeated steps needs (the host looks at the counters between
        those).  On return all sequences are in lockstep again."""
        if use_graph and self._graph is None:
            self.seq_steps.copy_(self.step_ctr.expand(self.B)); self.in_fb.zero_()
            self.target.copy_(self.step_ctr + 1)
            self.capture()
        self.seq_steps.copy_(self.step_ctr.expand(self.B))
        self.in_fb.zero_()
        self.target.copy_(self.step_ctr + n_steps)

        def run(k):
            for _ in range(k):
                if use_graph:
                    self._graph.replay()
                else:
                    self.step()
                self.launch_count += self.launches_per_step
                self.n_spec_steps += 1

        run(n_steps)
        while not bool((self.seq_steps >= self.target).all()):
            run(1 if self.B <= 8 else 2)
        self.step_ctr.add_(n_steps)
        self._seq_steps_done += n_steps * self.B
        self._last_src = "mixed"

