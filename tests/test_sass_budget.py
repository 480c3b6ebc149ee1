"""Instruction budget of the shipped throughput kernels, checked on the built library's SASS (CPU only).

The block kernels are bound by register-file operand delivery and their unrolled loop bodies just
fit the SM's instruction cache, so three things must not creep in unnoticed
with a source or toolchain change: local-memory traffic on the main path, a loop body that outgrows
the instruction cache, and extra operand-delivery cycles.  profiles/static_rf.py applies the model
of profiles/rf_model.py to `cuobjdump -sass`; the bounds below are the shipped build's numbers
(embed 1,818 instructions / 2,589 cycles, extract 1,042 / 1,359 per 32 blocks) plus ~2 %.
"""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "secure-video-steganography-using-ecc-and-dct_b200", "libsvs_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    spec = importlib.util.spec_from_file_location("static_rf", os.path.join(ROOT, "profiles", "static_rf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.functions(LIB)


def _one(fns, key):
    names = [n for n in fns if key in n]
    assert len(names) == 1, names
    return fns[names[0]]


@pytest.mark.parametrize("key,max_instr,max_cycles,fp32_cycles", [
    # blk::embed_blk_kernel<3,1,true,false>: BGR in, gray stego out, all 63 coefficients
    ("embed_blk_kernelILi3ELi1ELb1ELb0", 1860, 2650, 2177),
    # blk::extract_blk_kernel<1,4>: gray stego in, 4 coefficient row pairs
    ("extract_blk_kernelILi1ELi4", 1065, 1390, 980),
])
def test_main_loop_budget(sass, key, max_instr, max_cycles, fp32_cycles):
    mod, fns = sass
    body = mod.loop_body(_one(fns, key))
    ops = [t.split()[1] if t.startswith("@") else t.split()[0] for _, t in body]
    spills = [o for o in ops if o.split(".")[0] in ("STL", "LDL")]
    assert not spills, "local-memory traffic on the main path: %s" % spills[:4]
    assert not any(o.startswith("CALL") for o in ops), "the rare repair call must stay off the main path"
    cycles, fp32, _, per_n = mod.model(body)
    assert sum(per_n.values()) <= max_instr
    assert cycles <= max_cycles
    # the op-exact arithmetic itself: 54 single-rounded operations per 8-point transform, no more, no fewer
    assert abs(fp32 - fp32_cycles) <= 8, fp32


def test_every_block_kernel_is_one_cta_of_16_warps(sass):
    """128 registers x 512 threads = the whole register file: one CTA per SM, by construction."""
    import re
    import subprocess
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
            name = None
    blk = {k: v for k, v in regs.items() if "3blk" in k}
    assert len(blk) == 24, sorted(blk)
    assert max(blk.values()) <= 128, {k: v for k, v in blk.items() if v > 128}
