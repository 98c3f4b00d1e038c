"""Host-side helpers on the input side of the training path — mirrors ``holocron.utils`` (data collation)."""
from . import data  # noqa: F401
