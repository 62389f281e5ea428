"""The device atan2f (ripcurrents_b200/csrc/atan2f_ref.h, used by vectorToColor) must return libm's bits: the header is
compiled for the host (same single-rounding fp32 operations, no FMA contraction) and compared with this image's glibc."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_atan2f_port_matches_libm(tmp_path):
    exe = str(tmp_path / "atan2f_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "ripcurrents_b200", "csrc"),
                           os.path.join(ROOT, "tests", "csrc", "atan2f_check.c"), "-o", exe, "-lm"])
    out = subprocess.check_output([exe, "20000000"], timeout=300).decode().split()
    assert out == ["0", "0"], out
