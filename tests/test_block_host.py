"""CPU tests of the block-level code of the packed kernels (csrc/svs_block.cuh).

The kernels are written on the machine operations of csrc/svs_hw.cuh, each of which also has a
plain C++ body; tests/host_math compiles block_embed / block_extract for the host, so the
register layouts (column pairs / row pairs), the scalar stage-1 regrouping, the division-free
quantiser with its repair path, the always-exact tie-prone coefficients and the bit packing are
checked against the oracle here, without a GPU.  (The -m gpu tests then check the same code as
compiled for sm_100a.)"""
import numpy as np
import pytest

from oracle import dctqim_oracle as onp
from tests.host_math import build as hm


def _blocks_of(img, nb):
    if img.ndim == 3:
        return np.ascontiguousarray(img.reshape(8, nb, 8, 3).transpose(1, 0, 2, 3))
    return np.ascontiguousarray(img.reshape(8, nb, 8).transpose(1, 0, 2))


def _run(L, rng, ch, delta, n, nb, lo=0, hi=256, saturated=False):
    """One strip of nb blocks (8 x 8nb pixels) through the oracle and through the host build of
    the kernels' block code; returns (#stego px, #gray px, #bits from stego, #bits from cover) differing."""
    shape = (8, 8 * nb, 3) if ch == 3 else (8, 8 * nb)
    img = rng.integers(lo, hi, shape, dtype=np.uint8)
    if saturated:
        img[:, : 8 * (nb // 4)] = 0
        img[:, 8 * (nb // 4): 8 * (nb // 2)] = 255
    bits = rng.integers(0, 2, nb * n, dtype=np.uint8)
    gray, stego, k = onp.embed_frame(img, delta, bits, n)
    assert k == nb * n
    px = _blocks_of(img, nb)
    st = np.zeros((nb, 8, 8), np.uint8)
    gr = np.zeros((nb, 8, 8), np.uint8)
    rc = L.hm_blk_embed(ch, px.ctypes.data, nb, float(delta), n, bits.ctypes.data, st.ctypes.data, gr.ctypes.data)
    assert rc == 0, "packed quantiser does not cover delta=%r" % delta
    d_stego = int((_blocks_of(stego, nb) != st).sum())
    d_gray = int((_blocks_of(gray, nb) != gr).sum())
    out = np.zeros(nb * n, np.uint8)
    sp = _blocks_of(stego, nb)
    assert L.hm_blk_extract(1, sp.ctypes.data, nb, float(delta), n, out.ctypes.data) == 0
    d_bits = int((onp.extract_frame_bits(stego, delta, n)[: nb * n] != out).sum())
    out2 = np.zeros(nb * n, np.uint8)
    assert L.hm_blk_extract(ch, px.ctypes.data, nb, float(delta), n, out2.ctypes.data) == 0
    d_cover = int((onp.extract_frame_bits(img, delta, n)[: nb * n] != out2).sum())
    return d_stego, d_gray, d_bits, d_cover


@pytest.mark.parametrize("ch", [1, 3])
@pytest.mark.parametrize("delta", [20, 1, 2.5, 3, 7, 8, 16, 100, 0.5])
def test_block_code_matches_oracle_for_every_coefficient_count(ch, delta):
    L = hm.load()
    rng = np.random.default_rng(int(delta * 8) + ch)
    for n in (63, 62, 48, 47, 33, 32, 31, 17, 16, 15, 10, 1):
        assert _run(L, rng, ch, delta, n, 160) == (0, 0, 0, 0), (ch, delta, n)


def test_block_code_repair_paths_on_many_blocks():
    """Enough blocks that the flagged-fraction repair (about 6e-5 per coefficient at delta 20)
    and the exact ties of coefficients (0,4), (4,0), (4,4) (one block in 160) occur many times."""
    L = hm.load()
    rng = np.random.default_rng(7)
    assert _run(L, rng, 3, 20, 63, 30000) == (0, 0, 0, 0)
    assert _run(L, rng, 1, 20, 63, 20000, 64, 192) == (0, 0, 0, 0)
    assert _run(L, rng, 1, 3, 63, 20000) == (0, 0, 0, 0)
    assert _run(L, rng, 3, 20, 10, 20000) == (0, 0, 0, 0)


def test_block_code_saturated_blocks():
    """All-black / all-white blocks: clipping on the way out and the reference's own wrong bits."""
    L = hm.load()
    rng = np.random.default_rng(3)
    for delta, n in ((20, 63), (100, 63), (8, 10), (1, 63)):
        assert _run(L, rng, 3, delta, n, 400, saturated=True) == (0, 0, 0, 0), (delta, n)
