"""Shape guards with the reference's names and AssertionError behaviour (dmm/utils/checker.py:4-41).

The hot path's only error convention in the reference is an AssertionError raised by these helpers
(SURVEY.md section 4); callers never catch it, so the drop-in keeps it."""


def _rank(t, want):
    got = len(t.shape)
    assert got == want, "get {} {}".format(tuple(t.shape), got)
    return t.shape


def CHECK2D(t):
    return _rank(t, 2)


def CHECK3D(t):
    return _rank(t, 3)


def CHECK4D(t):
    return _rank(t, 4)


def CHECK5D(t):
    return _rank(t, 5)


def CHECKEQ(a, b, s=None):
    assert a == b, "get {} {}".format(a, b)


def CHECKSIZE(t, size):
    want = tuple(size.shape) if hasattr(size, "shape") else tuple(size)
    assert tuple(t.shape) == want, "get {} {}".format(tuple(t.shape), want)
