"""The shipped binary still polls.

The persistent kernel's cross-GPU gather polls peer-written memory with WEAK loads (they overlap; strong loads of one warp
serialise).  ptxas may treat a weak load of an unchanged address as loop-invariant and once compiled the polling loop into a
single pass (profiles/r02_fused_exchange_ab.md) -- a silent change that dead-locks a shared-step group.  The source defeats it
by offsetting every polling round's addresses with `clock64() >> 63`; this test reads the SASS of the built library and
checks that each such clock read sits inside a loop: a later branch jumps back to an address at or before it.  No GPU needed.
"""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "tfdiffeq_b200", "libb2ode.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

KERNELS = [
    "_Z16k_fused_adaptiveId9RhsLorenzIdELi7ELi512EEv11FusedParams",       # BASELINE config 2 (dopri5, fp64)
    "_Z16k_fused_adaptiveIf9RhsLorenzIfELi7ELi512EEv11FusedParams",
    "_Z16k_fused_adaptiveId16RhsLotkaVolterraIdELi4ELi512EEv11FusedParams",
]


def _sass(kernel):
    out = subprocess.run([CUOBJDUMP, "-sass", "-fun", kernel, LIB], capture_output=True, text=True, timeout=300).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


@pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")
@pytest.mark.parametrize("kernel", KERNELS)
def test_weak_polling_loops_survive_ptxas(kernel):
    ins = _sass(kernel)
    assert ins, "kernel not found in the library: " + kernel
    clocks = [a for a, t in ins if "SR_CLOCKLO" in t]
    # three exchanges are inlined in the comm warp (two of the initial-step heuristic, one per attempt), each with a batched
    # polling round and a tail loop for grids beyond 160 blocks
    assert len(clocks) >= 6, "expected the polling rounds' clock reads, found %d" % len(clocks)
    branches = [(a, int(m.group(1), 16)) for a, t in ins for m in [re.search(r"\bBRA\b.*\b0x([0-9a-f]+)\s*$", t)] if m]
    for c in clocks:
        assert any(a > c and tgt <= c and a - c < 0x1000 for a, tgt in branches), \
            "the polling round at 0x%x is no longer inside a loop: ptxas removed the poll" % c
    # and the loads it polls with are still there, next to the clock read
    for c in clocks:
        near = [t for a, t in ins if c < a < c + 0x400]
        assert any(t.split()[0].startswith("LDG") or " LDG" in t for t in near), "no load after the clock read at 0x%x" % c
