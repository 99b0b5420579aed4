"""On-disk formats of the loop: `{iter:06d}-rho.npz`
checkpoints with key `rho_design_elements`, `histories.npz`, the exported config,
and the restart path of `initialize_density` (reference
common_density.py:313-346,718-734)."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_checkpoints_and_restart():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sktopt
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=4, record_times=4)
        opt = sktopt.core.OC_Optimizer(cfg, sktopt.mesh.toy_problem.toy_test())
        opt.parameterize()
        opt.optimize()
        design = np.asarray(opt.tsk.design_elements)
        rho_end = opt._state.rho.cpu().numpy()
        comp = np.asarray(opt.recorder.as_object().compliance)
        assert comp.size == 4
        # one checkpoint per export tick (every iteration here), reference key
        for it in (1, 2, 3, 4):
            path = os.path.join(tmp, "data", f"{it:06d}-rho.npz")
            assert os.path.exists(path), path
            with np.load(path) as d:
                assert d.files == ["rho_design_elements"]
                assert d["rho_design_elements"].shape == design.shape
        with np.load(os.path.join(tmp, "data", "000004-rho.npz")) as d:
            assert np.array_equal(d["rho_design_elements"], rho_end[design])
        with np.load(os.path.join(tmp, "data", "000002-rho.npz")) as d:
            rho2 = d["rho_design_elements"].copy()
        with np.load(os.path.join(tmp, "histories.npz"), allow_pickle=True) as h:
            assert "compliance" in h.files and "vol_error" in h.files
            assert np.allclose(np.ravel(h["compliance"]), comp)
        assert sktopt.core.misc.find_latest_iter_file(os.path.join(tmp, "data"))[0] == 4

        # restart from iteration 2: densities come from the checkpoint, the loop
        # resumes at iteration 3 and runs to max_iters
        cfg2 = sktopt.core.OC_Config(dst_path=tmp, max_iters=4, record_times=4,
                                     restart=True, restart_from=2)
        opt2 = sktopt.core.OC_Optimizer(cfg2, sktopt.mesh.toy_problem.toy_test())
        opt2.parameterize()
        opt2._ensure_state_initialized()
        st = opt2._state
        assert st.iter_begin == 3 and st.iter_end == 5
        assert np.array_equal(st.rho.cpu().numpy()[design], rho2)
        opt2.optimize()
        comp2 = np.asarray(opt2.recorder.as_object().compliance)
        assert comp2.size == 4 + 2 and np.all(np.isfinite(comp2))   # imported history + 2 new
        # the density field itself carries the state the OC update needs except the
        # running sensitivity scale: the resumed iterates stay close to the original run
        assert np.max(np.abs(comp2[-2:] - comp[-2:]) / comp[-2:]) <= 5e-2

        # newest checkpoint when restart_from is not given
        cfg3 = sktopt.core.OC_Config(dst_path=tmp, max_iters=4, record_times=4, restart=True)
        opt3 = sktopt.core.OC_Optimizer(cfg3, sktopt.mesh.toy_problem.toy_test())
        opt3.parameterize()
        opt3._ensure_state_initialized()
        assert opt3._state.iter_begin == 5          # nothing left to do
