"""SAMS generator (reference: models/networks/sams/): SPADE / MultiSpade / AttentiveMultiSpade residual blocks in an
encoder - middle - decoder stack.  Same classes, constructor arguments and state_dict keys as the reference; the forward
pass drives the hand-written kernels (tcgen05 convs + csrc/spade.cu) and is an inference engine (eval mode)."""
from .attentive_multispade import AttentiveMultiSpade
from .multispade import MultiSpade
from .sams_generator import SamsGenerator
from .spade import SPADE, AnySpadeResBlock, SynchronizedBatchNorm2d

__all__ = ["SPADE", "MultiSpade", "AttentiveMultiSpade", "AnySpadeResBlock", "SamsGenerator", "SynchronizedBatchNorm2d"]
