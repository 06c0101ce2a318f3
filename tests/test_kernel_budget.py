"""Static budgets of the granule kernel that its speed depends on, checked on the built library and with ptxas
here, without a GPU:
  * no instance of the granule kernel spills -- local memory misses L1 there (the shared-memory carve-out is at its maximum),
    three spilled loads per granule measured 9 % of the kernel's time;
  * the address span of the granule loop of the benchmarked instance stays well inside what was measured to work with the
    32 KB instruction cache (cold code belongs in __noinline__ functions, outside the loop's range)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
KERNEL = "_ZN3l3b17l3_granule_kernelILi2ELi4ELb0ELb0ELb0ELb0EEEvNS_11BatchParamsEPKNS_4TileEj"

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")


def test_no_granule_kernel_instance_spills(tmp_path):
    from audio_formats_b200 import build
    src = str(build.CSRC / "l3_kernels.cu")
    res = subprocess.run(["nvcc", *[f for f in build.NVCC_FLAGS if f != "-shared"], "-Xptxas", "-v", "-c", "-o", str(tmp_path / "k.o"), src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stderr.split("\n")
    seen = 0
    for i, l in enumerate(lines):
        if "Function properties for" in l and "l3_granule_kernel" in l:
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", lines[i + 1])
            assert m, lines[i + 1]
            assert m.group(1) == "0" and m.group(2) == "0", (l, lines[i + 1])
            seen += 1
    assert seen >= 18, seen   # stereo / mono x arithmetic modes x delivery formats x Layer III / Layer I-II, + the tap instances


def test_granule_loop_span_fits_the_instruction_cache(built):
    import audio_formats_b200 as af
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, str(af.library_path())], capture_output=True, text=True).stdout
    ins = re.findall(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", sass, re.M)
    assert len(ins) > 3000, "kernel not found in the library"
    span = 0
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.U)?\s+(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < int(addr, 16):
            span = max(span, int(addr, 16) - int(m.group(1), 16))
    # measured: 2,987 instructions (47 KB) -> 17.6 ms with luck in the layout, 2,222 -> 17.3 ms whatever the layout; 16 bytes each
    assert 0 < span // 16 <= 2400, span // 16


def test_alias_reduction_neighbour_reads_stay_inside_the_warp_buffer():
    """Long-block granules read the neighbouring bands' eight elements straight from the spectrum buffer (l3_kernels.cu,
    "alias reduction across all 31 band boundaries"): lane 0 reads eight elements below the spectrum, lane 31 eight above its
    576 coefficients, both dropped afterwards.  Those reads must stay inside the warp's own Dbuf and be 16-byte aligned for
    the stereo instance (float2 elements, float4 loads).  Pins the layout constants the kernel relies on."""
    src = (ROOT / "audio_formats_b200" / "csrc" / "l3_kernels.cu").read_text()
    hdr = (ROOT / "audio_formats_b200" / "csrc" / "l3_kernels.cuh").read_text()
    stride = int(re.search(r"constexpr int kDStride = (\d+);", src).group(1))
    xr_len = int(re.search(r"constexpr int kXrStride = (\d+);", hdr).group(1))
    assert re.search(r"Dbuf\[1 \+ 15 \* kDStride \+ kXrStride\]", src), "WarpSmem::Dbuf layout changed"
    xr0 = 1 + 15 * stride                      # index of xr[0] inside Dbuf (D = Dbuf + 1, xr = D + 15 rows)
    total = 1 + 15 * stride + xr_len
    for lane in range(32):
        dn = xr0 + (lane - 1) * 18 + 10        # band-1, elements 10..17
        up = xr0 + (lane + 1) * 18             # band+1, elements 0..7
        assert 0 <= dn and dn + 8 <= total, (lane, dn)
        assert 0 <= up and up + 8 <= total, (lane, up)
        assert (dn * 8) % 16 == 0 and (up * 8) % 16 == 0, (lane, dn, up)   # float2 elements, Dbuf is 16-byte aligned
