"""Command-line options with the reference's flag names and defaults (options/base_options.py, train_options.py,
test_options.py): a two-phase parse in which the chosen model class adds / overrides its own flags."""
from .base_options import BaseOptions  # noqa: F401
from .test_options import TestOptions  # noqa: F401
from .train_options import TrainOptions  # noqa: F401
