from .core import *  # noqa: F401,F403
from .trainers import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
