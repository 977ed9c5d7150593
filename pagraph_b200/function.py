"""Builtin message / reduce descriptors, mirroring the `dgl.function` names the reference models use
(`fn.copy_src`, `fn.sum`, `fn.mean` — PaGraph/model/gcn_nssc.py:71-74, graphsage_nssc.py:98-106)."""


class CopySrc:
    def __init__(self, src, out):
        self.src, self.out = src, out


class Reduce:
    def __init__(self, mode, msg, out):
        self.mode, self.msg, self.out = mode, msg, out


def copy_src(src, out):
    return CopySrc(src, out)


def copy_u(u, out):
    return CopySrc(u, out)


def sum(msg, out):  # noqa: A001 - dgl's name
    return Reduce("sum", msg, out)


def mean(msg, out):
    return Reduce("mean", msg, out)


def max(msg, out):  # noqa: A001
    raise NotImplementedError("fn.max (GraphSAGE 'pool') is outside the rebuilt hot path; "
                              "the reference trainers use 'mean' (examples/profile/pa_gs.py)")
