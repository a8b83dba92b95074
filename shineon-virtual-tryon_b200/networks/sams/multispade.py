"""MultiSpade: one SPADE per label map, applied one after the other in sorted key order
(reference: models/networks/sams/multispade.py:8-77)."""
from torch import Tensor, nn

from .spade import SPADE


class MultiSpade(SPADE):
    DEFAULT_KEY = "default_key"

    def __init__(self, config_text, norm_nc, label_channels_dict, activation, sort_fn=sorted):
        nn.Module.__init__(self)  # duck-types SPADE like the reference: none of SPADE's own layers are created
        if isinstance(label_channels_dict, int):
            label_channels_dict = {MultiSpade.DEFAULT_KEY: label_channels_dict}
        self.sort_fn = sort_fn
        self.label_channels = label_channels_dict
        self.spade_layers = nn.ModuleDict({key: SPADE(config_text, norm_nc, label_nc, activation)
                                           for key, label_nc in label_channels_dict.items()})

    def try_fix_labelmap_dict(self, tensor):
        if len(self.spade_layers) == 1:
            return {list(self.spade_layers.keys())[0]: tensor}
        raise ValueError("You passed a single Tensor, but I don't know which spade layer to pass it through. "
                         f"My spade layers are:\n{self.spade_layers}.")

    def _ordered(self, labelmap_dict):
        if isinstance(labelmap_dict, Tensor):
            labelmap_dict = self.try_fix_labelmap_dict(labelmap_dict)
        assert len(labelmap_dict) == len(self.spade_layers), f"{len(labelmap_dict)=} != {len(self.spade_layers)=}"
        return list(self.sort_fn(labelmap_dict.items()))

    def run(self, ctx, x, labelmap_dict, *, act=None, act_param=0.0, shared=None, **out):
        """multispade.py:48-65: every layer but the last hands an un-activated f32 tensor to the next one (which
        normalises it again); the last applies the caller's activation and output selection."""
        items = self._ordered(labelmap_dict)
        for key, seg in items[:-1]:
            x, _ = self.spade_layers[key].run(ctx, x, seg, shared=shared, want_f32=True, want_planes=False)
        key, seg = items[-1]
        return self.spade_layers[key].run(ctx, x, seg, act=act, act_param=act_param, shared=shared, **out)
