from . import classification, detection, segmentation, utils  # noqa: F401
from .classification import *  # noqa: F401,F403
from .detection import *  # noqa: F401,F403
from .segmentation import *  # noqa: F401,F403
