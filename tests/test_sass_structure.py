"""The built library carries what DESIGN.md section 4 says the SpMV kernel is made of -- checked on the SASS of the
in-tree libtilespmv_b200.so (no GPU needed; cuobjdump comes with the toolkit):

  UBLKCP   1-D TMA bulk copy global -> shared (the per-warp chunk ring)
  SYNCS    mbarrier arrive / try_wait (completion of the bulk copies)
  LDGSTS   cp.async staging of the x operand
  UBLKPF   bulk L2 prefetch of the x window (gather-bound launches)
  PREEXIT  griddepcontrol.launch_dependents, ACQBULK  griddepcontrol.wait (programmatic dependent launch)

and only sm_100a code, no tensor-core instructions (SpMV has no dense contraction) and no local-memory spills."""
import re
import shutil
import subprocess

import pytest

from tilespmv_b200 import _capi

cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(shutil.which(cuobjdump) is None, reason="cuobjdump not available")


@pytest.fixture(scope="module")
def kernels():
    _capi.load()  # builds the library if the sources changed (nvcc cross-compiles without a GPU)
    txt = subprocess.run([cuobjdump, "-sass", _capi.lib_path()], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, _, body = part.partition("\n")
        out[name.strip()] = body
    return txt, out


def test_only_sm_100a_images(kernels):
    txt, _ = kernels
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("prec", ["d", "f"])
@pytest.mark.parametrize("plain", [0, 1])
def test_spmv_kernel_uses_tma_mbarrier_cp_async_and_pdl(kernels, prec, plain):
    _, k = kernels
    name = f"_ZN3tsp16tile_spmv_kernelI{prec}Li2ELi96ELb{plain}EEEvNS_8SpmvArgsIT_EE"
    assert name in k, [n for n in k if "tile_spmv" in n]
    body = k[name]
    for mnemonic in ("UBLKCP", "SYNCS", "LDGSTS", "UBLKPF", "PREEXIT", "ACQBULK"):
        assert re.search(r"\b" + mnemonic + r"\b", body), mnemonic
    assert not re.search(r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCMMA|HGMMA)\b", body)
    assert not re.search(r"\b(STL|LDL)\b", body), "local-memory spill in the hot kernel"
    # the grid-dependency release is the first thing the kernel does: before any shared- or global-memory access
    first = re.search(r"\b(PREEXIT|LDS|STS|LDG|STG|UBLKCP|LDGSTS)\b", body)
    assert first and first.group(1) == "PREEXIT"


def test_plain_epilogue_is_shorter(kernels):
    _, k = kernels
    n = {p: len(re.findall(r"/\*[0-9a-f]{4}\*/", k[f"_ZN3tsp16tile_spmv_kernelIdLi2ELi96ELb{p}EEEvNS_8SpmvArgsIT_EE"])) for p in (0, 1)}
    assert n[1] < n[0], n
