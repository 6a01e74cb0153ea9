"""Minimal CGNS/Python tree accessors used by the chimera path.

Re-statement (not a copy) of the handful of Converter.Internal functions the
reference's setInterpData / setInterpTransfers Python layer calls
(/root/reference/Cassiopee/Converter/Converter/Internal.py: getValue :2086,
setValue :297, copyRef :2514, getZones :1678, createChild/_createChild,
getNodeFromName1/2, getNodesFromType1/2).  A node is the plain list
``[name, value, children, type]``; values are numpy arrays (strings are 'c'
char arrays, integers are E_NpyInt=int32, reals float64, Fortran order).
"""
import numpy

E_NpyInt = numpy.int32

__GridCoordinates__ = 'GridCoordinates'
__FlowSolutionNodes__ = 'FlowSolution'
__FlowSolutionCenters__ = 'FlowSolution#Centers'


def isStdNode(node):
    """-1: a single standard node, 0: a list of standard nodes, -2: other."""
    if not isinstance(node, list) or len(node) == 0:
        return -2
    if len(node) == 4 and isinstance(node[0], str) and isinstance(node[2], list) and isinstance(node[3], str):
        return -1
    n0 = node[0]
    if isinstance(n0, list) and len(n0) == 4 and isinstance(n0[0], str):
        return 0
    return -2


def setValue(node, value=None):
    if value is None:
        node[1] = None
    elif isinstance(value, numpy.ndarray):
        node[1] = value if value.flags.f_contiguous else numpy.asfortranarray(value)
    elif isinstance(value, (int, numpy.integer)):
        node[1] = numpy.array([value], dtype=E_NpyInt)
    elif isinstance(value, (float, numpy.floating)):
        node[1] = numpy.array([value], dtype=numpy.float64)
    elif isinstance(value, str):
        node[1] = numpy.array([c for c in value], 'c')
    elif isinstance(value, (list, tuple)):
        t = value
        while isinstance(t, (list, tuple)):
            t = t[0]
        if isinstance(t, (float, numpy.floating)):
            node[1] = numpy.array(value, dtype=numpy.float64, order='F')
        elif isinstance(t, (int, numpy.integer)):
            node[1] = numpy.array(value, dtype=E_NpyInt, order='F')
        else:
            raise TypeError("setValue: unsupported list value.")
    else:
        node[1] = numpy.array([value])


def getValue(node):
    n = node[1]
    if isinstance(n, numpy.ndarray):
        if n.dtype.char in ('S', 'c'):
            if n.ndim == 1:
                return n.tobytes().decode()
            return [n[:, i].tobytes().decode().strip() for i in range(n.shape[1])]
        if n.dtype in (numpy.int32, numpy.int64):
            return int(n.flat[0]) if n.size == 1 else n
        if n.dtype == numpy.float64:
            return float(n.flat[0]) if n.size == 1 else n
    return n


def createNode(name, ntype, value=None, children=None, parent=None):
    node = [name, None, [] if children is None else children, ntype]
    setValue(node, value)
    if parent is not None:
        parent[2].append(node)
    return node


def createChild(node, name, ntype, value=None, children=None, pos=-1):
    c = createNode(name, ntype, value, children)
    if pos == -1:
        node[2].append(c)
    else:
        node[2].insert(pos, c)
    return c


def _createChild(node, name, ntype, value=None, children=None, pos=-1):
    createChild(node, name, ntype, value, children, pos)
    return None


def createUniqueChild(node, name, ntype, value=None, children=None, pos=-1):
    c = getNodeFromName1(node, name)
    if c is None:
        return createChild(node, name, ntype, value, children, pos)
    c[3] = ntype
    setValue(c, value)
    if children is not None:
        c[2] = children
    return c


def getNodeFromName1(node, name):
    for c in node[2]:
        if c[0] == name:
            return c
    return None


def getNodesFromName1(node, name):
    return [c for c in node[2] if c[0] == name]


def getNodeFromName2(node, name):
    for c in node[2]:
        if c[0] == name:
            return c
    for c in node[2]:
        for d in c[2]:
            if d[0] == name:
                return d
    return None


def getNodeFromName(node, name):
    if node[0] == name:
        return node
    for c in node[2]:
        r = getNodeFromName(c, name)
        if r is not None:
            return r
    return None


def getNodesFromType1(node, ntype):
    return [c for c in node[2] if c[3] == ntype]


def getNodeFromType1(node, ntype):
    for c in node[2]:
        if c[3] == ntype:
            return c
    return None


def _getNodesFromType2(node, ntype, result):
    if node[3] == ntype:
        result.append(node)
        return
    for c in node[2]:
        if c[3] == ntype:
            result.append(c)
        else:
            for d in c[2]:
                if d[3] == ntype:
                    result.append(d)


def getNodesFromType2(node, ntype):
    result = []
    std = isStdNode(node)
    if std == 0:
        for c in node:
            _getNodesFromType2(c, ntype, result)
    elif std == -1:
        _getNodesFromType2(node, ntype, result)
    return result


def getZones(t):
    """All Zone_t nodes of a tree, a base, a zone or a list of zones."""
    return getNodesFromType2(t, 'Zone_t')


def getBases(t):
    if isStdNode(t) == -1 and t[3] == 'CGNSTree_t':
        return getNodesFromType1(t, 'CGNSBase_t')
    return []


def _dup(node):
    return [node[0], node[1], [_dup(c) for c in node[2]], node[3]]


def copyRef(node):
    """Copy of the tree structure sharing the numpy values (Internal.py:2514)."""
    std = isStdNode(node)
    if std == -1:
        return _dup(node)
    if std == 0:
        return [_dup(n) for n in node]
    return node


def _rmNodesByName1(node, name):
    node[2][:] = [c for c in node[2] if c[0] != name]
    return None


def getZoneDim(zone):
    """['Structured', ni, nj, nk, celldim] for a structured zone."""
    v = zone[1]
    if v.shape[0] == 3:
        return ['Structured', int(v[0, 0]), int(v[1, 0]), int(v[2, 0]), 3]
    if v.shape[0] == 2:
        return ['Structured', int(v[0, 0]), int(v[1, 0]), 1, 2]
    return ['Structured', int(v[0, 0]), 1, 1, 1]


def newCGNSTree():
    t = ['CGNSTree', None, [], 'CGNSTree_t']
    createChild(t, 'CGNSLibraryVersion', 'CGNSLibraryVersion_t', None)
    t[2][0][1] = numpy.array([3.1], dtype=numpy.float32)
    return t


def newCGNSBase(name='Base', cellDim=3, physDim=3, parent=None):
    b = [name, numpy.array([cellDim, physDim], dtype=E_NpyInt), [], 'CGNSBase_t']
    if parent is not None:
        parent[2].append(b)
    return b


def newZone(name, ni, nj, nk, parent=None):
    """Structured Zone_t; value rows i,j,k = [npts, ncells, 0]."""
    v = numpy.zeros((3, 3), dtype=E_NpyInt, order='F')
    v[0, 0] = ni; v[1, 0] = nj; v[2, 0] = nk
    v[0, 1] = max(ni - 1, 1); v[1, 1] = max(nj - 1, 1); v[2, 1] = max(nk - 1, 1)
    z = [name, v, [], 'Zone_t']
    createChild(z, 'ZoneType', 'ZoneType_t', 'Structured')
    if parent is not None:
        parent[2].append(z)
    return z
