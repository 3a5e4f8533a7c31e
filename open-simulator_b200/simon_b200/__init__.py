"""simon_b200 — host side of the B200-native pod->node placement engine (mirrors pkg/simulator of the reference)."""
from .objects import AppResource, ResourceTypes  # noqa: F401
