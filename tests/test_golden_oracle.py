"""The oracle must reproduce the committed fixtures bit for bit (CPU)."""
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["blobs33", "wave48x4", "rect40x24_f6d2", "stiff33"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    row, col = int(z["row"]), int(z["col"])
    over = {}
    for k, v in zip(z["over_keys"], z["over_vals"]):
        over[str(k)] = int(v) if str(k) in ("alpha", "beta", "fSelect", "dSelect", "gSelect") else float(v)
    return z, row, col, over


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    ol.build_oracle()
    z, row, col, over = load_case(name)
    o = ol.Oracle(row, col, ol.default_params(row, **over))
    o.set_state(z["M0"], z["C0"])
    for k in range(int(z["steps"])):
        r = o.step()
        assert r["countIters"] == z["countIters"][k]
        assert r["nits"] == [int(v) for v in z["nits"][k][:len(r["nits"])]]
        M, C = o.get_state()
        assert np.array_equal(M, z[f"M{k + 1}"])
        assert np.array_equal(C, z[f"C{k + 1}"])
