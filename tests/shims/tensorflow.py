"""Bare stand-in: the reference's tests import tensorflow without using it (tests/test_controllers.py:4)."""
