"""import-only stub"""
from . import animation  # noqa
rcParams = {}


def use(*a, **k):
    pass
