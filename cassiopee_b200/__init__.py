"""cassiopee_b200 — B200-native chimera interpolation hot path of Cassiopee's
Connector (setInterpData / setInterpTransfers, storage='inverse' layout),
hand-written sm_100a CUDA behind the reference's Python API.

    import cassiopee_b200.Connector.PyTree as X
    import cassiopee_b200.Converter as C
    import cassiopee_b200.Generator as G
    import cassiopee_b200.Internal as Internal

The CUDA library (csrc/libcassiopee_b200.so) is loaded on first use; there is
no CPU fallback.
"""
__version__ = "0.1.0"
