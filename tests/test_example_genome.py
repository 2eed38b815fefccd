"""examples/genome_distmix.py (window-sharded run through the ternary-row chromosome driver) against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from gauss_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_genome_example_matches_oracle(oracle):
    spec = importlib.util.spec_from_file_location("genome_distmix", os.path.join(ROOT, "examples", "genome_distmix.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run(n_windows=6, per_mb=400.0)
    windows, res = out["windows"], out["results"]
    assert sorted(res) == list(range(len(windows)))
    full = synth.make_genotypes(len(out["bp"]), out["sizes"], seed=out["seed"])
    for wi in (0, len(windows) - 1):
        x = windows[wi]
        assert res[wi][2] == 0
        idx = np.sort(np.concatenate([x["measured"], x["unmeasured"]]))
        r = oracle.run_window(out["type"][idx], out["bp"][idx], out["z_site"][idx], full[idx], out["sizes"], out["w"],
                              x["start_bp"], x["end_bp"])
        um = np.isin(idx, x["unmeasured"])
        assert np.abs(r["z"][um] - res[wi][0]).max() <= 1e-6
        assert np.abs(r["info"][um] - res[wi][1]).max() <= 1e-6
