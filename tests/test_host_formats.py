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
