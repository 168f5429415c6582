from dropoutdecoding_b200.models.llama import *  # noqa: F401,F403
from dropoutdecoding_b200.models.llama import CustomLlamaForCausalLM  # noqa: F401
