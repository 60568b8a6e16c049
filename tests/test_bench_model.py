"""bench.py's byte and update model against the figures of SURVEY.md 8(d) (CPU, no GPU needed)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_level_sizes():
    assert bench.level_sizes(1025) == [1025, 513, 257, 129, 65, 33, 17, 9, 5, 3]
    assert len(bench.level_sizes(2049)) == 11 and len(bench.level_sizes(257)) == 8


def test_algorithmic_bytes_match_survey():
    # 3D fp64 V(2,2): 257^3 = 2.547 GB, 1025^3 = 161.32 GB, 2049^3 = 1288.3 GB; fp32 1025^3 = 80.66 GB
    assert abs(bench.algorithmic_bytes_per_cycle(257, 8) / 1e9 - 2.547) < 0.001
    assert abs(bench.algorithmic_bytes_per_cycle(1025, 8) / 1e9 - 161.32) < 0.01
    assert abs(bench.algorithmic_bytes_per_cycle(2049, 8) / 1e9 - 1288.3) < 0.1
    assert abs(bench.algorithmic_bytes_per_cycle(1025, 4) / 1e9 - 80.66) < 0.01


def test_updates_match_survey():
    # (nu1+nu2) * sum_l (n_l-2)^3: 257^3 -> 7.565e7, 1025^3 -> 4.892e9, 2049^3 -> 3.920e10
    assert abs(bench.updates_per_cycle(257) / 7.565e7 - 1) < 1e-3
    assert abs(bench.updates_per_cycle(1025) / 4.892e9 - 1) < 1e-3
    assert abs(bench.updates_per_cycle(2049) / 3.920e10 - 1) < 1e-3


def test_measured_peak_source():
    peak, src = bench.measured_peak_gbs()
    assert peak > 1000 and isinstance(src, str)
