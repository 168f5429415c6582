from dropoutdecoding_b200.models.instructblip import *  # noqa: F401,F403
from dropoutdecoding_b200.models.instructblip import (CustomInstructBlipForConditionalGeneration, CustomLlamaForCausalLM,  # noqa: F401
                                                      select_by_vote, select_by_average, seed)
