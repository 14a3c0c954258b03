"""Import-only stand-in for matplotlib (absent from this image) used by the drop-in tests: the reference's Examples/OC
scripts do ``import matplotlib.pyplot as plt`` at the top and never plot inside their learning loops.  Test infrastructure;
never on the product path (JinEnv imports matplotlib lazily, only inside play_animation)."""


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Anything()
