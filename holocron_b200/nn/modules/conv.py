"""placeholder (filled in later this round)."""
__all__ = []
