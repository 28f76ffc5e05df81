"""ORBmatcher / Frame view bindings (filled in with the matcher kernels)."""


class FrameView:
    pass


class ORBmatcher:
    def __init__(self, *a, **k):
        raise NotImplementedError("matcher kernels are not built yet")
