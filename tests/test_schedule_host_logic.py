"""Host side of the one-sequence speculative schedule (dropoutdecoding_b200/schedules.py::_spec_steps) on a scripted
fake engine: which pieces run in which order, how many steps a call commits, when the lookahead is allowed, and the
EOS stop -- no GPU involved.  (What the pieces compute is covered bit for bit by tests/test_engine_gpu.py.)"""
import pytest
import torch

from dropoutdecoding_b200.schedules import SpeculativeSchedules


class _Stream:
    def synchronize(self):
        pass


class FakeEngine(SpeculativeSchedules):
    """fb[s]: step s needs its masked members repeated; ok[s]: the guess made for step s is confirmed;
    tok[s]: the token step s produces (tok[-1]... the token the call starts from is `first`)."""

    def __init__(self, fb, ok, tok, first=7, lookahead=True):
        self.fb, self.ok, self.tok, self.first = fb, ok, tok, first
        self.spec_active, self.lookahead, self.speculative = True, lookahead, "auto"
        self.need_fb_host, self.chain_flags_host, self.tok_host, self.tok2_host = [0], [0, 0], [0], [0]
        self.n_spec_steps = self.n_fallbacks = self.n_chain = self.n_chain_ok = 0
        self._spec_base = (0, 0)
        self.launch_count = self.launches_per_step = 0
        self.done, self.log = 0, []          # committed steps, pieces in launch order

    def _cur(self):
        return self.first if self.done == 0 else self.tok[self.done - 1]

    def _run_piece(self, graphs, name, fn):
        self.log.append(name)
        if name == "head":
            self.tok_host[0] = self._cur()
            self.need_fb_host[0] = int(self.fb[self.done])
        elif name == "chain":
            self.tok_host[0] = self._cur()
            t = self.done
            self.done += 1
            self.tok2_host[0] = self.tok[t]          # the token the pass's own step produced
            nxt = self.fb[t + 1] if t + 1 < len(self.fb) else 0
            self.chain_flags_host[:] = [int(self.ok[t]), int(nxt)]
        else:
            assert name in ("fin", "fin_chain", "fallback")
            self.done += 1

    # never reached by these scripts
    _spec_head = _spec_fin = _spec_chain = _spec_fin_chain = _spec_fallback = None


@pytest.fixture(autouse=True)
def no_cuda_stream(monkeypatch):
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())


def run(fb, ok=None, tok=None, n=None, **kw):
    n = len(fb) if n is None else n
    eng = FakeEngine(fb, ok or [1] * len(fb), tok or list(range(100, 100 + len(fb))), **kw)
    return eng, eng._spec_steps(n, None, None)


def test_no_repeats_is_head_fin_per_step():
    eng, done = run([0, 0, 0, 0])
    assert done == eng.done == 4 and eng.log == ["head", "fin"] * 4
    assert eng.n_spec_steps == 4 and eng.n_fallbacks == eng.n_chain == 0


def test_confirmed_lookahead_finishes_two_steps_from_one_pass():
    eng, done = run([0, 1, 0, 0])
    assert done == eng.done == 4
    assert eng.log == ["head", "fin", "head", "chain", "fin_chain", "head", "fin"]
    assert (eng.n_fallbacks, eng.n_chain, eng.n_chain_ok) == (1, 1, 1)


def test_confirmed_lookahead_into_another_repeat_chains_without_a_head():
    eng, done = run([1, 1, 1, 0, 0])
    assert done == eng.done == 5
    assert eng.log == ["head", "chain", "chain", "chain", "fin_chain", "head", "fin"]
    assert (eng.n_fallbacks, eng.n_chain, eng.n_chain_ok) == (3, 3, 3)


def test_wrong_guess_drops_the_lookahead():
    eng, done = run([0, 1, 0, 0], ok=[1, 0, 1, 1])
    assert done == eng.done == 4
    assert eng.log == ["head", "fin", "head", "chain", "head", "fin", "head", "fin"]
    assert (eng.n_chain, eng.n_chain_ok) == (1, 0)


def test_last_step_of_a_call_uses_the_plain_repeat():
    eng, done = run([0, 0, 1])
    assert done == eng.done == 3 and eng.log == ["head", "fin", "head", "fin", "head", "fallback"]
    eng, done = run([1, 1], ok=[1, 1])          # the lookahead of step 0 covers step 1, whose repeat is the last step
    assert done == eng.done == 2 and eng.log == ["head", "chain", "fallback"]


def test_without_lookahead_every_repeat_is_a_second_pass():
    eng, done = run([1, 0, 1, 0], lookahead=False)
    assert done == 4 and eng.log == ["head", "fallback", "head", "fin", "head", "fallback", "head", "fin"]


def test_eos_stops_within_one_pass():
    fb, tok = [0, 0, 0, 0, 0, 0], [100, 101, 2, 103, 104, 105]           # step 2 produces EOS (= 2)
    eng = FakeEngine(fb, [1] * 6, tok)
    done = eng._spec_steps(6, None, stop_token=2)
    assert done == eng.done == 3 and eng.log == ["head", "fin"] * 3 + ["head"]      # the 4th head only sees the EOS
    # EOS out of the step a lookahead pass completes: the chain stops right there
    fb, tok = [1, 1, 1, 0, 0, 0], [2, 101, 102, 103, 104, 105]           # step 0 produces EOS inside a chain
    eng = FakeEngine(fb, [1] * 6, tok)
    done = eng._spec_steps(6, None, stop_token=2)
    assert done == eng.done == 1 and eng.log == ["head", "chain"]
    # ... also when the guess was confirmed and the following step needs no repeat (chain -> fin_chain): the lookahead
    # rows are NOT finished past the EOS (round-1 bug: decoding ran on to max_new_tokens)
    fb, tok = [1, 0, 0, 0, 0, 0], [2, 101, 102, 103, 104, 105]
    eng = FakeEngine(fb, [1] * 6, tok)
    done = eng._spec_steps(6, None, stop_token=2)
    assert done == eng.done == 1 and eng.log == ["head", "chain"]
    # EOS out of the step finished by fin_chain is seen by the next head
    fb, tok = [1, 0, 0, 0, 0, 0], [100, 2, 102, 103, 104, 105]
    eng = FakeEngine(fb, [1] * 6, tok)
    done = eng._spec_steps(6, None, stop_token=2)
    assert done == eng.done == 2 and eng.log == ["head", "chain", "fin_chain", "head"]
    # the call starts on EOS: nothing is decoded
    eng = FakeEngine([0, 0], [1, 1], [100, 101], first=2)
    assert eng._spec_steps(2, None, stop_token=2) == 0 and eng.done == 0 and eng.log == ["head"]
