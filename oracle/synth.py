"""Synthetic assets live in the product package (tepose_b200/synthetic.py); re-exported for the oracle / tests."""
from tepose_b200.synthetic import *  # noqa: F401,F403
from tepose_b200.synthetic import _rng, _uniform  # noqa: F401
