"""`models/llama.py` of the reference is a verbatim copy of transformers-4.44's
LlamaForCausalLM.forward as `CustomLlamaForCausalLM` (model -> lm_head -> .float(),
models/llama.py:27-47), imported by nothing.  It is kept in the API surface: the class below is
the installed transformers LlamaForCausalLM, whose forward computes exactly that."""
from transformers import LlamaForCausalLM


class CustomLlamaForCausalLM(LlamaForCausalLM):
    pass
