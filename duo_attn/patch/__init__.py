from duo_attention_b200.patch import *  # noqa: F401,F403
from duo_attention_b200.patch import __all__  # noqa: F401
