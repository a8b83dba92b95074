"""models/networks/attention/__init__.py: the attention layers selectable by name."""
from . import sagan

ATTENTION_TYPES = {"sagan": sagan.SelfAttention}
