"""tilespmv_b200 -- B200-native (sm_100a) TileSpMV behind the reference's C-level surface.

The product is the C-ABI shared library ``libtilespmv_b200.so`` declared in ``include/tilespmv.h``;
this package is only the thin host-side mirror used by the tests and bench.py.
"""
__version__ = "0.1.0"
