"""Connector.PyTree public names of the chimera path
(/root/reference/Cassiopee/Connector/Connector/PyTree.py:7-13 re-exports the
same names from OversetData.py)."""
from .OversetData import (setInterpData, _setInterpData, setInterpData2, _setInterpData2, _setInterpDataChimera, setInterpTransfers,
                          _setInterpTransfers, setInterpTransfersD, _setInterpTransfersD,
                          _createInterpRegion__, _addCellN__, applyBCOverlaps, _applyBCOverlaps,
                          getIntersectingDomains, setHoleInterpolatedPoints, _setHoleInterpolatedPoints)
from .Connector import setInterpData__, getInterpolatedPoints__, createHooks__
from .compactTransfers import ___setInterpTransfers, miseAPlatDonorTree__

__all__ = ['setInterpData', '_setInterpData', 'setInterpData2', '_setInterpData2', 'setInterpTransfers', '_setInterpTransfers',
           'setInterpTransfersD', '_setInterpTransfersD', 'applyBCOverlaps', '_applyBCOverlaps',
           'getIntersectingDomains', 'setHoleInterpolatedPoints', '_setHoleInterpolatedPoints',
           'miseAPlatDonorTree__', '___setInterpTransfers']
