"""CPU checks of the entropy-stage device logic (host build of csrc/bj_entropy.cuh, orchestrated like
the kernels) against the reference-generated golden coefficient planes: bit-exact after every scan."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names
from entropy_sim import decode_file


@pytest.mark.parametrize("name", golden_case_names())
@pytest.mark.parametrize("sub_bits", [1024, 128])
def test_entropy_logic_matches_reference(name, sub_bits):
    data = (GOLDEN / "cases" / f"{name}.jpg").read_bytes()
    z = np.load(GOLDEN / "cases" / f"{name}.npz")
    p, per_scan, _ = decode_file(data, sub_bits=sub_bits)
    if p.progressive:
        for k, grids in enumerate(per_scan, start=1):
            for c, g in enumerate(grids):
                assert np.array_equal(g, z[f"scan{k}_coef{c}"]), (k, c)
    for c, g in enumerate(per_scan[-1]):
        assert np.array_equal(g, z[f"coef{c}"]), c


def test_entropy_logic_base_image(golden_meta):
    import hashlib
    data = (GOLDEN / "base_image.jpg").read_bytes()
    p, per_scan, stats = decode_file(data)
    for k, grids in enumerate(per_scan):
        got = [hashlib.sha256(np.ascontiguousarray(g).tobytes()).hexdigest() for g in grids]
        assert got == golden_meta["base_image"]["scan_coef_sha256"][k], k
