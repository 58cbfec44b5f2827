"""CPU: the wire formats either side of the hot path: Case CSV ingest (costmap.py:134-156)
and the Solution_*.csv trajectory writer (animation/record_solution.py:22-51, pandas to_csv with sep='\\t')."""
import filecmp
import os

import numpy as np

from automatedvaletparking_b200 import scenarios as scn
from automatedvaletparking_b200.animation.record_solution import DataRecorder, COLUMNS
from conftest import GOLDEN


def test_trajectory_writer_matches_pandas(tmp_path):
    pd = __import__("pytest").importorskip("pandas")
    rng = np.random.default_rng(0)
    traj = rng.normal(size=(64, 8)) * np.array([10, 10, 1, 1, 1, 0.1, 0.1, 5])
    traj[3, 2] = 0.0; traj[4, 3] = 1e-7; traj[5, 0] = 123456789.125; traj[6, 1] = -0.0; traj[7, 4] = 1e22; traj[8, 5] = 5e-324
    traj = [list(map(float, r)) for r in traj]
    DataRecorder.record(str(tmp_path), "A.csv", traj)
    df = pd.DataFrame(traj)
    df.columns = COLUMNS
    df.to_csv(str(tmp_path / "ref.csv"), index='True', sep='\t')              # record_solution.py:36-51 verbatim
    assert filecmp.cmp(str(tmp_path / "Solution_A.csv"), str(tmp_path / "ref.csv"), shallow=False)


def test_stored_solution_round_trips_byte_identically(tmp_path):
    """first rows of the reference's stored solution/Solution_Case1.csv (fixture): read -> record == original bytes"""
    src = os.path.join(GOLDEN, "solution_case1_head.tsv")
    a = DataRecorder.read(src)
    assert a.shape == (11, 8)
    assert a[0, 0] == -16.0199004975124 and a[0, 1] == -13.5074626865672        # starts at Case1's start pose
    DataRecorder.record(str(tmp_path), "B.csv", [list(r) for r in a])
    assert open(str(tmp_path / "Solution_B.csv")).read() == open(src).read()


def test_record_rejects_wrong_width(tmp_path):
    import pytest
    with pytest.raises(AssertionError):
        DataRecorder.record(str(tmp_path), "C.csv", [[0.0] * 7])


def test_case_csv_ingest_round_trip(tmp_path):
    """all 20 BenchmarkCases: row -> Scenario -> row is the identity on the doubles; ragged obstacle lists, |theta| > pi kept raw"""
    for c in range(1, 21):
        s = scn.benchmark_case(c)
        p = str(tmp_path / f"Case{c}.csv")
        scn.write_case_csv(s, p)
        t = scn.read_case_csv(p)
        assert scn.case_row(t) == scn.case_row(s)
        assert len(t.obs) == len(s.obs) and all(np.array_equal(a, b) for a, b in zip(t.obs, s.obs))
    s10 = scn.benchmark_case(10)
    assert abs(s10.theta0) > np.pi                                                   # Appendix A: un-normalised headings survive ingest
    empty = scn.Scenario(0.0, 0.0, 0.1, 6.0, 3.0, 0.4, [], None, "empty")
    p = str(tmp_path / "empty.csv")
    scn.write_case_csv(empty, p)
    assert scn.read_case_csv(p).obs == []
