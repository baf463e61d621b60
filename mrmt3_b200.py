"""Alias so that `import mrmt3_b200` resolves to the hyphenated package directory `mr-mt3_b200/`."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("mr-mt3_b200")
