"""CPU oracle for the Dropout-Decoding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dropoutdecoding_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the CPU baseline, never as the product path.

Pinning status: the reference ships no golden vectors and no tests
(SURVEY.md section 4), so the pure functions restated in ``dropout_oracle.py``
are pinned against the reference's *own code* instead: ``tests/golden/
make_golden.py`` imports the reference's functions from ``/root/reference``
in the build container, checks the restatement against them bit for bit and
writes the fixtures committed under ``tests/golden/``.  The decoder forward
(transformers==4.44.0, absent from the container) is restated from the
published Llama algorithm and pinned against the installed transformers
``LlamaForCausalLM`` (fp32, eager) by ``tests/test_oracle_lm.py``.
"""
from .dropout_oracle import *  # noqa: F401,F403
