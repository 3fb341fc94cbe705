"""Parses include/niw_b200.h so that tests and ``build()`` can check the built library against the
declared C ABI."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "niw_b200.h")


def declared_symbols(path=HEADER):
    """Names of every function declared in the public header."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    return sorted(set(re.findall(r"\b(niw_[a-z0-9_]+)\s*\(", text)))
