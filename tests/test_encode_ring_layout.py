"""CPU check of the shared-memory layout encode1024_ring_kernel (csrc/encode.cu) relies on: the XOR swizzle is a
permutation of the 256 chunks of a row, its three access patterns are bank-conflict free, and the split
(lane part) ^ (compile-time part) + immediate address formulas written into the kernel equal chunk ^ swz(chunk >> 3)."""
import os
import runpy


def test_ring_swizzle_identities(capsys):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "profiles", "debug", "check_ring_swizzle.py"), run_name="__main__")
    assert "split formulas exact" in capsys.readouterr().out
