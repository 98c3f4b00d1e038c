from .repvgg import *  # noqa: F401,F403
from .darknet import *  # noqa: F401,F403
from .rexnet import *  # noqa: F401,F403
from .resnet import *  # noqa: F401,F403
from .mobileone import *  # noqa: F401,F403
from .res2net import *  # noqa: F401,F403
from .sknet import *  # noqa: F401,F403
from .convnext import *  # noqa: F401,F403
