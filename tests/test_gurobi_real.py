"""With a REAL gurobipy (absent from the build container; the size-limited licence allows 2,000 variables / 2,000 constraints): the
144-cell saved run of the reference (examples/simulated_st: 1,152 pairs + 144 + 144 + triangles < 2,000 variables) solved through
both model builders on the GPU path.  The reference's saved solution matches every aligned cell to the reference cell with the
same index (examples/simulated_st/matches_df.csv); with identical candidates, costs and model order Gurobi must
return it again.  Skipped when gurobipy cannot be imported or has no usable licence."""
import numpy as np
import pytest

from tests.util import golden_frame, golden_params, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["gurobi", "gurobi_rows"])
def test_simulated_st_known_answer_with_real_gurobi(backend, tmp_path, monkeypatch):
    gp = pytest.importorskip("gurobipy")
    if not hasattr(gp, "Env") or not hasattr(getattr(gp, "GRB", None), "METHOD_PDHG"):
        pytest.skip("gurobipy >= 13 needed (GRB.METHOD_PDHG, src/same.py:1169-1170)")
    try:
        gp.Env(params={"OutputFlag": 0}).dispose()
    except Exception as e:       # no licence on this machine
        pytest.skip(f"no usable Gurobi licence: {e}")
    import same_b200
    monkeypatch.chdir(tmp_path)
    g = load_golden("simulated_st")
    ref_df, al_df = golden_frame(g, "ref"), golden_frame(g, "aligned")
    ct = [str(c) for c in g["commonCT"]]
    matches, var_out = same_b200.run_same(ref_df, al_df, ct, outprefix=None, optim_params=golden_params(g, "optim"),
                                          gurobi_params=golden_params(g, "gurobi"), solver=backend)
    assert len(var_out["x"]) == 1152 and len(matches) == 144
    assert np.array_equal(matches["aligned_idx"].to_numpy(), matches["ref_idx"].to_numpy())
