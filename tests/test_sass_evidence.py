"""CPU-side look at the machine code of the shipped library (cuobjdump, no GPU): the default <= 64-marker kernel of the bench
configuration is Blackwell tensor-core code -- tcgen05.mma (UTCHMMA), tensor-memory loads / stores (LDTM / STTM), bulk async
copies (UBLKCP), packed FP32 (FFMA2) -- and runs without register spills (no local-memory instructions)."""
import re
import shutil
import subprocess

import pytest

KERNEL = "_ZN4bann6k1_tc5ILi5ELi5ELi1ELi0ELb1ELi7ELb1EEEvNS_6K1ArgsE"     # k1_tc5<5,5,1,tanh,LEAN,NCT=7,DEFER>


def test_k1_tc5_is_tcgen05_code_without_spills():
    import rs_bann_b200 as rb
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, rb.LIB_PATH], capture_output=True, text=True).stdout
    assert "Function : " + KERNEL in out, "k1_tc5<5,5,1,tanh> is not in the library"
    count = lambda op: len(re.findall(r"\b" + op + r"\b", out))
    assert count("UTCHMMA") >= 24          # 8 forward + 16 backward MMAs of a super-tile (plus the drain)
    assert count(r"STTM\.x4") >= 14 and count(r"LDTM\.x16") >= 2 and count("UBLKCP") >= 2
    assert count("FFMA2") >= 80 and count(r"MUFU\.EX2") == 20 and count(r"MUFU\.RCP") >= 20
    assert count("LDL") == 0 and count("STL") == 0, "register spills in the hot kernel"
