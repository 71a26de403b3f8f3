"""Test suite: `-m "not gpu"` runs on a CPU box, `-m gpu` on a B200 (see conftest.py)."""
