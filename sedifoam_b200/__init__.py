"""sedifoam_b200 -- B200-native particle hot path of sediFoam behind the reference's own C-ABI.

The product is ``libsedi_b200.so`` (CUDA, sm_100a; sources in ``csrc/``, C-ABI in ``include/sedi_b200.h``).  This
package is the thin ctypes binding used by the tests and by ``bench.py``; it mirrors the reference's boundary
(``interfaceToLammps/library.h:29-63``) one function per method.  There is no CPU fallback: importing works anywhere
(so the symbol table can be checked), but any compute call aborts without a CUDA device, and a missing library
raises immediately.
"""
from .lib import Lammps, build_library, library_path, load_library, EXPORTED_SYMBOLS  # noqa: F401
from .lib import decomp_grid, decomp_owner, decomp_links  # noqa: F401
from .lib import DRAG_ERGUN_WENYU, DRAG_SYAMLAL_OBRIEN, FORCE_DRAG, FORCE_PGRAD, FORCE_BUOY, FORCE_ADDEDMASS, FORCE_LIFT  # noqa: F401
from .lib import FORCE_HISTORY, FORCE_WALL_LUB, FORCE_INLET  # noqa: F401

__all__ = ["Lammps", "build_library", "library_path", "load_library", "EXPORTED_SYMBOLS"]
