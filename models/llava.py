from dropoutdecoding_b200.models.llava import *  # noqa: F401,F403
from dropoutdecoding_b200.models.llava import (CustomLlavaForConditionalGeneration, select_by_vote,  # noqa: F401
                                               select_by_average, seed)
