"""Stub of matplotlib.pyplot; see matplotlib/__init__.py in this directory."""


def __getattr__(name):  # any plotting call in the reference is off the hot path
    raise RuntimeError("matplotlib stub: plotting (%s) is not available in the oracle harness" % name)
