"""CPU tests of the on-disk formats the loop writes and reads back (host code,
no CUDA): cfg.json round trip (reference common_density.py:313-346) and the
recorder's histories.npz (tools/history.py:534-687)."""
import os
import tempfile

import numpy as np


def test_config_json_round_trip():
    import sktopt
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.LogMOC_Config(
            dst_path=tmp, max_iters=40, record_times=8, solver_option="cg_pyamg",
            vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3))
        cfg.export(tmp)
        assert os.path.exists(os.path.join(tmp, "cfg.json"))
        back = sktopt.core.LogMOC_Config.import_from(tmp)
        assert type(back) is type(cfg)
        assert back.max_iters == 40 and back.record_times == 8
        assert back.solver_option == "cg_pyamg" and back.solver_config.solver == cfg.solver_config.solver
        assert back.vol_frac.target_value == 0.3 and back.vol_frac.scheduler_type == cfg.vol_frac.scheduler_type
        assert back.p.target_value == cfg.p.target_value and back.beta.curvature == cfg.beta.curvature
        assert back.mu_p == cfg.mu_p and back.lagrangian_percentile == cfg.lagrangian_percentile


def test_histories_npz_round_trip_and_append():
    import sktopt
    with tempfile.TemporaryDirectory() as tmp:
        rec = sktopt.tools.HistoryCollection(tmp)
        rec.add("compliance", ylog=True)
        rec.add("vol_error")
        rec.add("dL", plot_type="min-max-mean-std")
        for i in range(4):
            rec.feed_data("compliance", 1.0 / (i + 1))
            rec.feed_data("vol_error", 0.1 * i)
            rec.feed_data("dL", np.arange(5.0) + i)
        rec.export_histories("histories.npz")
        with np.load(os.path.join(tmp, "histories.npz"), allow_pickle=True) as h:
            assert {"compliance", "compliance_header", "vol_error", "dL", "dL_header"} <= set(h.files)
            assert np.allclose(h["compliance"], [1.0, 0.5, 1.0 / 3.0, 0.25])
            assert h["dL"].shape == (4, 4)                    # min, mean, max, std x 4 records
            assert np.allclose(h["dL"][0], [0.0, 1.0, 2.0, 3.0])
        rec2 = sktopt.tools.HistoryCollection(tmp)
        rec2.add("compliance", ylog=True)
        rec2.add("vol_error")
        rec2.add("dL", plot_type="min-max-mean-std")
        rec2.import_histories()
        rec2.feed_data("compliance", 0.123)                   # a resumed run appends
        assert np.allclose(rec2.as_object().compliance, [1.0, 0.5, 1.0 / 3.0, 0.25, 0.123])
        assert rec2.latest("vol_error") == 0.1 * 3


def test_vtu_writer_round_trip(tmp_path):
    """info_mesh-*.vtu / condition.vtu (reference core/visualization.py:22-84,
    mesh/task_common.py:360-396) without meshio: fields, connectivity in the
    mesh's own local order and VTK cell types survive a write/read cycle."""
    import numpy as np
    from sktopt._fem import MeshHex, MeshTet
    from sktopt.core.visualization import export_mesh_with_info, read_vtu
    ax = [np.linspace(0, 1, n) for n in (4, 3, 3)]
    rng = np.random.default_rng(0)
    for mesh, vtk_type in ((MeshHex.init_tensor(*ax), 12), (MeshTet.init_tensor(*ax), 10)):
        rho = rng.uniform(size=mesh.t.shape[1])
        en = rng.uniform(size=mesh.t.shape[1])
        col = rng.integers(0, 4, mesh.p.shape[1])
        f = str(tmp_path / f"m{vtk_type}.vtu")
        export_mesh_with_info(mesh, point_data_values=[col], point_data_names=["node_color"],
                              cell_data_values=[rho, en],
                              cell_data_names=["rho_projected", "energy"], filepath=f)
        r = read_vtu(f)
        assert np.array_equal(r["points"], mesh.p.T)
        assert np.array_equal(r["connectivity"].reshape(-1, mesh.t.shape[0]), mesh.t.T)
        assert np.all(r["types"] == vtk_type)
        assert np.array_equal(r["offsets"], mesh.t.shape[0] * np.arange(1, mesh.t.shape[1] + 1))
        assert np.array_equal(r["cell_data"]["rho_projected"], rho)
        assert np.array_equal(r["cell_data"]["energy"], en)
        assert np.array_equal(r["point_data"]["node_color"], col)


def test_task_scale_keeps_the_load_list_and_condition_vtu(tmp_path):
    """scale() must scale the load arrays in place, not through the
    ``force`` property (n_tasks went from 1 to n_dof)."""
    import numpy as np
    import sktopt
    from sktopt.core.visualization import read_vtu
    tsk = sktopt.mesh.toy_problem.toy_test()
    f0 = tsk.neumann_linear[0].copy()
    p0 = tsk.mesh.p.copy()
    assert tsk.n_tasks == 1
    tsk.scale(0.125, 1e-5)
    assert tsk.n_tasks == 1 and isinstance(tsk.neumann_linear, list)
    assert np.allclose(tsk.neumann_linear[0], f0 * 1e-5, rtol=1e-15, atol=0)
    assert np.allclose(tsk.mesh.p, p0 * 0.125)
    tsk.scale(8.0, 1e5)
    assert np.allclose(tsk.neumann_linear[0], f0, rtol=1e-14, atol=0)
    tsk.export_analysis_condition_on_mesh(str(tmp_path))
    r = read_vtu(str(tmp_path / "condition.vtu"))
    assert set(np.unique(r["cell_data"]["condition"])) <= {0, 1, 2, 3}
    assert r["point_data"]["node_color"].max() == 2
