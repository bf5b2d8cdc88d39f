"""Stand-in for `dgl`, which the reference's dataset.py imports only to construct an unused
`dgl.DGLGraph()` (GNNAdvisor/dataset.py:5,56).  dgl is not installable offline."""


class DGLGraph(object):
    def __init__(self, *args, **kwargs):
        pass
