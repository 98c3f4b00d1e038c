"""placeholder (filled in later this round)."""
def add2d(*a, **k):
    raise NotImplementedError("add2d: kernel not built yet")


def norm_conv2d(*a, **k):
    raise NotImplementedError("norm_conv2d: kernel not built yet")


