"""Records draw calls instead of rendering (see package docstring)."""
CALLS = []


class _Axes(object):
    def __getattr__(self, name):
        def _rec(*a, **k):
            CALLS.append(("ax." + name, a, k))
        return _rec


class _Figure(object):
    pass


def subplots(*a, **k):
    return _Figure(), _Axes()


def Rectangle(*a, **k):
    return ("Rectangle", a, k)


def _noop(*a, **k):
    CALLS.append(("plt", a, k))


axis = tight_layout = draw = show = figure = imshow = savefig = _noop
