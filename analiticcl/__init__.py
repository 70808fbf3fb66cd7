"""`import analiticcl` -- the module name of the reference's Python binding (bindings/python/src/lib.rs), served
by the B200-native implementation: the reference's own test (bindings/python/tests/tests.py) and example
(bindings/python/examples/example.py) run unchanged against this package.  Everything lives in analiticcl_b200."""
from analiticcl_b200 import SearchParameters, VariantModel, VocabParams, Weights  # noqa: F401

__all__ = ["VariantModel", "Weights", "SearchParameters", "VocabParams"]
