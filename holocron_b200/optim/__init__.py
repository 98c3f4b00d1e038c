from .adabelief import *  # noqa: F401,F403
from .adamp import *  # noqa: F401,F403
from .adan import *  # noqa: F401,F403
from .ademamix import *  # noqa: F401,F403
from .lamb import *  # noqa: F401,F403
from .lars import *  # noqa: F401,F403
from .ralars import *  # noqa: F401,F403
from .tadam import *  # noqa: F401,F403
from . import wrapper  # noqa: F401  (Lookahead lives in holocron.optim.wrapper like in the reference)
