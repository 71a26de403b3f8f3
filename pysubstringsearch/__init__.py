"""Drop-in import name: `import pysubstringsearch` resolves to the B200 implementation,
so code (and the reference's own tests/test_pysubstringsearch.py) written against
Intsights/PySubstringSearch runs unchanged."""
from pysubstringsearch_b200 import Reader, Writer, pysubstringsearch  # noqa: F401

__all__ = ["Writer", "Reader"]
