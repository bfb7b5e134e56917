from .pyplot import *  # noqa: F401,F403
from .pyplot import Figure, close, colorbar, subplots  # noqa: F401
