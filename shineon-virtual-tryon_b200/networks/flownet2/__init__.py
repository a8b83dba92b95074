"""FlowNet2 native-op mirrors (reference: models/flownet2_pytorch/networks/*_package)."""
from .native_ops import ChannelNorm, Correlation, Resample2d  # noqa: F401
