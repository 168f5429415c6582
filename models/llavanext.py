from dropoutdecoding_b200.models.llavanext import *  # noqa: F401,F403
from dropoutdecoding_b200.models.llavanext import CustomLlavaNextForConditionalGeneration, select_by_vote, argmax, seed  # noqa: F401
