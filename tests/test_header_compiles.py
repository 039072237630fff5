"""include/tilespmv.h is the drop-in boundary: it must be consumable by a plain C host (the reference is C
compiled as CUDA C++), by C++, with and without the reference's unsuffixed names, in both precisions -- and the
reference-names mode must give a caller exactly the identifiers src/main.cu uses."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
GCC = shutil.which("gcc") or "/usr/bin/gcc"
GPP = shutil.which("g++") or "/usr/bin/g++"

CALLER = r"""
#include <stddef.h>
#define TILESPMV_REFERENCE_NAMES
%s
#include "tilespmv.h"
/* the calls of src/main.cu:87-180 by their reference names */
int drive(int rowA, int colA, int nnzA, int *rp, int *ci, MAT_VAL_TYPE *v, MAT_VAL_TYPE *x, MAT_VAL_TYPE *y, char *name)
{
    Tile_matrix M;
    int rowblkblock = 0;
    unsigned int *idx = NULL;
    int *c0 = NULL, *c1 = NULL;
    Tile_create(&M, rowA, colA, nnzA, rp, ci, v);
    if (M.tilenum < 0)
        return 1;
    {
        int p1[1], p2[1];
        if (M.tilenum == 1 && tilespmv_prepare(&M, p1, p2, &rowblkblock, &idx, &c0, &c1, rowA) != TILESPMV_OK)
            return 2;
        call_tilespmv_cuda(name, &M, p1, p2, rowblkblock, idx, c0, c1, rowA, colA, nnzA, rp, ci, v, (MAT_VAL_TYPE)1, x, y, y);
    }
    Tile_destroy(&M);
    return BLOCK_SIZE == 16 && sizeof(M.Format[0]) == 1 ? 0 : 3;
}
"""


def _compile(compiler, std, source, tmp_path, name):
    src = tmp_path / name
    src.write_text(source)
    r = subprocess.run([compiler, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I" + INC, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("std", ["-std=c99", "-std=c11", "-std=gnu17"])
def test_header_is_plain_c(tmp_path, std):
    _compile(GCC, std, '#include "tilespmv.h"\nint main(void) { return sizeof(Tile_matrix_f64) == sizeof(Tile_matrix_f32) ? 0 : 1; }\n',
             tmp_path, "plain.c")


@pytest.mark.parametrize("std", ["-std=c++11", "-std=c++17"])
def test_header_is_cplusplus_with_c_linkage(tmp_path, std):
    _compile(GPP, std, '#include "tilespmv.h"\nint main() { return tilespmv_version() != nullptr ? 0 : 1; }\n', tmp_path, "plain.cpp")


@pytest.mark.parametrize("precision", ["", "#define MAT_VAL_TYPE float\n#define TILESPMV_USE_F32"])
def test_reference_names_drive_the_library_like_main_cu(tmp_path, precision):
    _compile(GCC, "-std=c99", CALLER % precision, tmp_path, "caller.c")
    _compile(GPP, "-std=c++17", CALLER % precision, tmp_path, "caller.cpp")
