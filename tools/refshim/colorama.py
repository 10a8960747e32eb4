def init(*a, **k):
    pass


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Style = _Codes()
Back = _Codes()
