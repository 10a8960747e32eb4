import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, reset=True, log_colors=None, style="%", **kw):
        fmt = (fmt or "%(message)s").replace("%(log_color)s", "").replace("%(reset)s", "")
        super().__init__(fmt, datefmt, style)
