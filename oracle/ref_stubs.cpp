/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Abort-stubs for two reference symbols that the compiled reference objects
 * reference but that are NOT on the structured-donor chimera path (Cartesian
 * element bounding boxes for CEBB tests, cylindrical-coordinate ADT). Defining
 * them here keeps the link closed without pulling in more of KCore. */
#include <cstdio>
#include <cstdlib>
#include "Fld/FldArray.h"
namespace K_COMPGEOM
{
void compCartEltsArray(E_Int, E_Int, E_Int, E_Int, E_Float, E_Float, E_Float,
                       E_Float, E_Float, E_Float, K_FLD::FldArray<E_Float>&,
                       K_FLD::FldArray<E_Float>&)
{ fprintf(stderr, "oracle stub: compCartEltsArray is out of scope\n"); abort(); }
}
namespace K_LOC
{
E_Int cart2Cyl(E_Int, E_Float*, E_Float*, E_Float*, E_Float, E_Float, E_Float,
               E_Float, E_Float, E_Float, E_Float*, E_Float*, E_Int, E_Int,
               E_Int, E_Int, E_Float)
{ fprintf(stderr, "oracle stub: cart2Cyl is out of scope\n"); abort(); }
}
