"""CPU test of the reference staging recipe (oracle/stage_ref.py): what is staged is byte-identical
to /root/reference, importable on its own, and behaves like the live reference."""
import os

import numpy as np
import pytest

from oracle import stage_ref
from tests.synth import synth_frames, synth_bits, bits_to_str

HAVE_SOURCE = os.path.isdir(stage_ref.SOURCE)


@pytest.mark.skipif(not (HAVE_SOURCE or stage_ref.available()), reason="neither /root/reference nor a staged copy")
def test_staged_reference_is_unmodified_and_importable():
    d = stage_ref.stage()
    assert d and stage_ref.available()
    if HAVE_SOURCE:
        for rel in stage_ref.FILES:
            assert open(os.path.join(d, rel), "rb").read() == open(os.path.join(stage_ref.SOURCE, rel), "rb").read(), rel
    ref = stage_ref.import_reference()
    assert os.path.dirname(ref["config_and_setup"].__file__) == stage_ref.DEST
    f = synth_frames("stage", (16, 24, 3))
    bits = synth_bits("stage", 60)
    from oracle import dctqim_oracle as onp
    g, s, n = ref["config_and_setup"].proses_frame_qim_dct(f, 'embed', 20, bits_to_str(bits), num_ac_coeffs_to_use=10)
    g2, s2, n2 = onp.embed_frame(f, 20, bits, 10)
    assert n == n2 and np.array_equal(g, g2) and np.array_equal(s, s2)


def test_a_tampered_copy_is_rejected(tmp_path, monkeypatch):
    if not stage_ref.available():
        pytest.skip("nothing staged")
    import shutil
    fake = tmp_path / "_ref"
    shutil.copytree(stage_ref.DEST, fake)
    with open(fake / "config_and_setup.py", "a") as f:
        f.write("\n# edited\n")
    monkeypatch.setattr(stage_ref, "DEST", str(fake))
    assert not stage_ref.available()
    with pytest.raises(RuntimeError):
        stage_ref.import_reference()
