"""No-op matplotlib stand-in (tools/demo.py:25 imports pyplot; drawing is out of scope for this build)."""


def use(*a, **k):
    pass
