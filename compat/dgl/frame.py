from pagraph_b200.nodeflow import Frame, FrameRef  # noqa: F401
