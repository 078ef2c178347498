"""TEST INFRASTRUCTURE ONLY (oracle shim) -- inert matplotlib so the reference imports."""


def use(*a, **k):
    pass


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())
