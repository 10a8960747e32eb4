"""import-only stub"""


class File:
    def __init__(self, *a, **k):
        raise ImportError("h5py is not available (stub)")
