"""TestOptions (options/test_options.py:4-32)."""
from .base_options import BaseOptions


class TestOptions(BaseOptions):
    def initialize(self, parser):
        parser = BaseOptions.initialize(self, parser)
        parser.add_argument("--no_shuffle", action="store_true", default=True)
        parser.set_defaults(datamode="test")
        self.is_train = False
        parser.add_argument("--result_dir", type=str, default="test_results", help="save test result outputs")
        parser.add_argument("--tryon_list", help="CSV of CLOTH_PATH, PERSON_ID pairs (reference dataset feature)")
        parser.add_argument("--random_tryon", action="store_true", help="Randomly choose cloth-person pairs for try-on.")
        return parser
