"""CPU: static checks on the SASS of the shipped sweep kernels (cuobjdump of the cross-compiled objects; no GPU needed).

The fast kernel's whole point is a dispatch frame in which the 16 accumulator registers are read and written IN PLACE by
every case.  Whether ptxas does that is fragile: moving a few cases around, or prefetching one more word in the frame, made
it fall back to two register sets with 15-32 IMAD.MOV copies per dispatch (seen twice while the kernel was written).
This test pins the property, and the presence of the Blackwell paths the design claims."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "enoki_b200", "csrc", "ek_sweep_fast.o")


def _sass(kernel_substr):
    if shutil.which("cuobjdump") is None or not os.path.exists(OBJ):
        pytest.skip("cuobjdump or ek_sweep_fast.o not available")
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    on, lines = False, []
    for l in out.splitlines():
        if "Function :" in l:
            on = kernel_substr in l
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                lines.append(m.group(2).strip())
    assert lines, "kernel not found in the object"
    return lines


@pytest.mark.parametrize("kernel", ["ek_fast_kernelILi256", "ek_fast_kernelILi128"])
def test_dispatch_frame_has_no_accumulator_copies(kernel):
    sass = _sass(kernel)
    # the body loop's fetch: the two 64-bit constant-bank loads of the instruction word, indexed by the program counter
    fetch = [i for i, l in enumerate(sass) if re.match(r"LDC\.64 R\d+, c\[0x0\]\[R\d+\+0x38[08]\]", l)]
    assert len(fetch) >= 2
    # the LAST such pair belongs to the hot loop (init / fini loops come with their own, earlier / later in the listing);
    # take the pair that is followed by the 8 predicated operand loads
    start = None
    for i in fetch:
        window = sass[i:i + 40]
        if sum(1 for l in window if re.match(r"@P\d LDS\.128", l)) >= 8:
            start = i
            break
    assert start is not None, "hot-loop fetch not found"
    frame = sass[max(start - 4, 0):start + 45]
    moves = [l for l in frame if re.match(r"(IMAD\.MOV\.U32|MOV) R\d+, (RZ, RZ, )?R\d+", l)]
    assert len(moves) <= 2, "ptxas copies the accumulator in the dispatch frame again:\n" + "\n".join(frame)
    # operand fetches use immediate group offsets (block size is a template constant)
    assert any(re.match(r"@P\d LDS\.128 R\d+, \[R\d+\+0x[0-9a-f]+\]", l) for l in frame)


def test_blackwell_paths_present():
    sass = _sass("ek_fast_kernelILi256")
    text = "\n".join(sass)
    assert "UBLKCP" in text, "TMA bulk copy (cp.async.bulk) missing"
    assert "UBLKPF" in text, "bulk L2 prefetch missing"
    assert "SYNCS" in text, "mbarrier instructions missing"
    assert text.count("FFMA2") > 100, "packed fp32 FMA missing"
    assert re.search(r"ATOMS\.ADD RZ", text), "fire-and-forget shared-memory reduction (red.shared.add) missing"
