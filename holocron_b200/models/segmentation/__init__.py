from .unet3p import *  # noqa: F401,F403
