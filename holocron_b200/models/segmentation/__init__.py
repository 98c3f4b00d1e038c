from .unet import *  # noqa: F401,F403
from .unet3p import *  # noqa: F401,F403
from .unetpp import *  # noqa: F401,F403
