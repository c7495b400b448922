"""CPU-side checks of bench.py's contract (no GPU): both arms describe the workload with the same `config` object, every
workload the docstring names is runnable by name, and the round-robin step accounting adds up to one full build."""
import importlib.util
import itertools
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("xr_bench", os.path.join(REPO, "bench.py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def test_config_is_shared_by_both_arms_and_names_the_workload(bench):
    for name, w in bench.WORKLOADS.items():
        cfg = bench.config_of(name)
        assert cfg == bench.config_of(name) and cfg["workload"] == name and "model" not in cfg
        assert "step" in cfg and "cache" in cfg
        if w["kind"] == "general":
            n_tri = len(list(itertools.combinations(range(w["n_frag"]), 3)))
            assert ("one of the %d trimer" % n_tri) in cfg["step"]
    assert {"cfg1", "cfg2", "cfg3", "cfg4", "cfg5"} <= set(bench.WORKLOADS)


def test_defaults_finish_within_minutes(bench, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    args = bench.parse_args()
    assert (args.gpus, args.impl, args.workload) == (1, "xr", "cfg4")
    assert args.warmup >= 3 and 1 <= args.steps <= 8            # 11.6 s per step at N = 1


def test_round_robin_flops_add_up_to_a_full_build(bench):
    """four consecutive steps = every dimer four times + every trimer once"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200 import synth
    system = synth.make_system("toy3", n_frag=4)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"])
    dimers = list(itertools.combinations(range(4), 2))
    trimers = list(itertools.combinations(range(4), 3))
    full, split = eng.algorithmic_flops(dimers, trimers)
    d_only, _ = eng.algorithmic_flops(dimers, ())
    per_step = [d_only + eng.algorithmic_flops((), [ms])[0] for ms in trimers]
    assert abs(sum(per_step) - (full + (len(trimers) - 1) * d_only)) <= 1e-9 * full
    assert abs(split["dimer"] - d_only) <= 1e-9 * d_only
