from dropoutdecoding_b200.models.config import *  # noqa: F401,F403
from dropoutdecoding_b200.models.config import settings  # noqa: F401
