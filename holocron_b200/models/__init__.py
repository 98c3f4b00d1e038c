from . import classification, utils  # noqa: F401
from .classification import *  # noqa: F401,F403
