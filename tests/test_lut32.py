"""The 32-bit Huffman LUT of the lane-decoupled entropy kernels, checked exhaustively on the host (CPU only):
tests/cpp/lut32_check.cpp pushes every code of every book through the kernels' lookup sequence."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("root_bits", [8, 9, 10])
def test_every_code_decodes_through_the_lut(tmp_path, root_bits):
    exe = tmp_path / f"lut32_check_{root_bits}"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-DL3B_HUFF_ROOT_BITS={root_bits}", "-I", str(ROOT / "audio_formats_b200" / "csrc"),
                           "-o", str(exe), str(ROOT / "tests" / "cpp" / "lut32_check.cpp")])
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "lut32 ok" in res.stdout and f"root bits {root_bits}" in res.stdout
