"""wr_repeat_add (csrc/repeat_add.cuh) must equal n sequential fp32 additions bit
for bit: 1.5 M random (x, s, n) cases incl. binade crossings, sign changes and
tie-prone steps, on the host build of the same header."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_repeat_add_bit_exact():
    exe = os.path.join(HERE, "_build", "repeat_add_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "repeat_add_check.cpp"), "-lm"],
                   check=True, cwd=HERE)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "bad 0" in out.stdout
