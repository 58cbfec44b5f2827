"""Drop-in for animation/record_solution.py: the trajectory wire format.

`DataRecorder.record(save_path, save_name, trajectory)` writes what the reference's pandas call
`DataFrame(trajectory, columns=[x, y, theta, v, a, sigma, omega, t]).to_csv(file, index='True', sep='\\t')`
writes (record_solution.py:22-51): a header line with an empty index cell, then one line per state with the
row number and repr-shortest floats, tab separated.  `read` parses such a file back (main.py --mode 1 uses
pandas.read_csv for that, main.py:161-169).  No pandas dependency."""
import os
from typing import List

import numpy as np

COLUMNS = ['x', 'y', 'theta', 'v', 'a', 'sigma', 'omega', 't']


def _fmt(v) -> str:
    """pandas' default float formatting for to_csv is repr(float) (shortest round-trip); ints stay ints"""
    if isinstance(v, (bool, np.bool_)):
        return str(bool(v))
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    f = float(v)
    if f != f:
        return ''                      # NaN -> empty field (na_rep default)
    if f in (float('inf'), float('-inf')):
        return 'inf' if f > 0 else '-inf'
    return repr(f)


class DataRecorder:
    def __init__(self) -> None:
        pass

    @staticmethod
    def record(save_path: str, save_name: str, trajectory: List[List]):
        assert len(trajectory[0]) == 8, 'the trajectory size should be 8'
        if not os.path.exists(save_path):
            os.makedirs(save_path)
        file_name = os.path.join(save_path, 'Solution_' + save_name)
        with open(file_name, 'w', newline='') as f:
            f.write('\t'.join([''] + COLUMNS) + '\n')
            for i, row in enumerate(trajectory):
                f.write('\t'.join([str(i)] + [_fmt(v) for v in row]) + '\n')
        return file_name

    @staticmethod
    def read(file_name: str) -> np.ndarray:
        """(n, 8) float64 array of a Solution_*.csv file"""
        rows = []
        with open(file_name, 'r') as f:
            header = f.readline().rstrip('\n').split('\t')
            assert header[1:] == COLUMNS, 'not a Solution_*.csv trajectory file'
            for line in f:
                parts = line.rstrip('\n').split('\t')
                if len(parts) == 9:
                    rows.append([float(p) if p != '' else float('nan') for p in parts[1:]])
        return np.array(rows, dtype=np.float64).reshape(-1, 8)

    @staticmethod
    def save_gif():
        pass
