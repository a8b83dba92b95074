"""Host-side mirrors of the reference nn.Modules on the hot path (same names, ctor args, state_dict keys)."""
