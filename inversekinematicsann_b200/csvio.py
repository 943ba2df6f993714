"""Array-native CSV I/O either side of the solver.

The reference CLI reads targets with ``pd.read_csv(path).values.tolist()`` (reference cli.py:242,272) and writes
results with ``pd.DataFrame(rows, columns=[...]).to_csv(path, index=False)`` (cli.py:45-49, 74-78): header line, no
index column, shortest round-trip decimal text.  At 1e7 rows that list-of-lists detour costs far more than the solve,
so these helpers go file <-> ndarray directly (Arrow's multi-threaded reader/writer) and hand the solver a contiguous
float32 array.  Files stay interchangeable with the reference's: same header names, same column order, every value
parses back to the same float (the text of a few values differs in form only, e.g. ``3`` for ``3.0``).
"""
import numpy as np

POINT_COLUMNS = ('x', 'y', 'z')
ANGLE_COLUMNS = ('theta1', 'theta2', 'theta3', 'theta4')


def read_csv_array(path, n_columns, dtype=np.float32, engine='arrow'):
    """Rows of a headed CSV as a C-contiguous (N, n_columns) array.

    ``engine='pandas'`` parses with the reference's own reader (whose default float parser is not correctly rounded:
    ~1 value in 1e4 differs from the exact parse in the last fp64 bit), ``'arrow'`` parses exactly and several times
    faster.  After the cast to float32 the two agree except on fp32 rounding ties.
    """
    if engine == 'pandas':
        import pandas as pd
        values = pd.read_csv(path).values
    elif engine == 'arrow':
        import pyarrow.csv as pacsv
        table = pacsv.read_csv(path)
        if table.num_columns != n_columns:
            raise ValueError(f'{path}: expected {n_columns} columns, found {table.num_columns}')
        values = np.empty((table.num_rows, n_columns), dtype=dtype)
        for i in range(n_columns):
            values[:, i] = table.column(i).to_numpy()
        return values
    else:
        raise ValueError(f'unknown CSV engine {engine!r}')
    if values.ndim != 2 or values.shape[1] != n_columns:
        raise ValueError(f'{path}: expected {n_columns} columns, found shape {values.shape}')
    return np.ascontiguousarray(values, dtype=dtype)


def write_csv_array(path, values, columns):
    """Write an (N, len(columns)) array the way the reference's ``to_csv(index=False)`` lays files out."""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    values = np.asarray(values)
    if values.ndim != 2 or values.shape[1] != len(columns):
        raise ValueError(f'expected shape (N, {len(columns)}), got {values.shape}')
    table = pa.table({name: np.ascontiguousarray(values[:, i]) for i, name in enumerate(columns)})
    with open(path, 'wb') as sink:  # Arrow always quotes header names; pandas (and the reference's files) do not
        sink.write((','.join(columns) + '\n').encode())
        pacsv.write_csv(table, sink, pacsv.WriteOptions(include_header=False, quoting_style='none'))


def read_points_csv(path, dtype=np.float32, engine='arrow'):
    """Targets file of the reference CLI (columns x,y,z) -> (N, 3) array."""
    return read_csv_array(path, 3, dtype, engine)


def write_points_csv(path, points):
    write_csv_array(path, points, POINT_COLUMNS)


def read_angles_csv(path, dtype=np.float32, engine='arrow'):
    return read_csv_array(path, 4, dtype, engine)


def write_angles_csv(path, angles):
    """Results file of the reference CLI (columns theta1..theta4)."""
    write_csv_array(path, angles, ANGLE_COLUMNS)


def solve_csv(ik_engine, points_csv, angles_csv=None, engine='arrow'):
    """``--inverse-kine --points in.csv --to-file out.csv`` without the list detour: read, ``ikine`` on the array,
    write.  ``ik_engine`` is a FabrikInverseKinematics / AnnInverseKinematics; exceptions propagate as in the
    reference (the CLI prints them, cli.py:250-252)."""
    points = read_points_csv(points_csv, engine=engine)
    angles = ik_engine.ikine(points, as_array=True)
    if angles_csv is not None:
        write_angles_csv(angles_csv, angles)
    return angles
