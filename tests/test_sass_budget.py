"""Static guards on the compiled tile kernels (cuobjdump on the in-tree object, no GPU needed): register budgets that
decide how many blocks an SM holds, no spills, no indirect branches in the pixel loops.  The three-component kernels once
sat at 128 registers (16 resident warps per SM) for a whole series of measurements before anybody looked
(profiles/r1_notes.md, r1i) -- this is the tripwire."""
import os
import re
import shutil
import subprocess

import pytest

from charls_b200 import build as product_build

OBJECT = os.path.join(product_build.OBJ_DIR, "jls_kernels.cu.o")

# registers per thread: 64 -> 32 resident one-warp blocks per SM, 72 -> 28, 80 -> 25
BUDGETS = {
    "k_decode_tiled": {1: 64, 2: 64, 3: 72, 4: 64},  # 3: the near-lossless ones take 72 since rows may end inside a word
    "k_encode_tiled": {1: 72, 2: 72, 3: 80, 4: 72},
}


@pytest.fixture(scope="module")
def resources(product):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    assert os.path.exists(OBJECT), OBJECT
    text = subprocess.run(["cuobjdump", "--dump-resource-usage", OBJECT], capture_output=True, text=True, check=True).stdout
    found = {}
    name = None
    for line in text.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            found[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    return found


def test_tile_kernels_stay_within_their_register_budgets(resources):
    seen = 0
    for mangled, (registers, stack) in resources.items():
        m = re.search(r"(k_(?:en|de)code_tiled)ILi(\d)E", mangled)
        if not m:
            continue
        seen += 1
        kernel, components = m.group(1), int(m.group(2))
        assert stack == 0, (mangled, "spills", stack)
        assert registers <= BUDGETS[kernel][components], (mangled, registers)
    # 2 kernels x 4 component counts x 8 / 16 bit containers x (near-lossless, lossless at any depth, lossless at full depth)
    assert seen == 48


def test_no_indirect_branch_in_the_tile_kernels(product):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    text = subprocess.run(["cuobjdump", "-sass", OBJECT], capture_output=True, text=True, check=True).stdout
    function, offenders = None, set()
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            function = m.group(1)
        elif function and "_tiled" in function and re.search(r"\bBRX\b", line):
            offenders.add(function)
    assert not offenders, offenders  # a jump table in a pixel loop (the colour transform once was one)
