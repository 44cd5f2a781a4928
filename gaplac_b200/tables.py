"""Table ingestion on the host side of the path: what the reference does with CSV.jl /
DataFrames before it calls the GP code - read a delimited table, keep complete cases, build the design matrix
`Matrix(df[!, vars])` (CLI/src/mcmc.jl:17-26, CLI/src/select.jl:38-47), rank-based inverse-normal transform of a
response (src/utils.jl:16-28), and packing many response columns into the y-batch of gpl_lml_batched (config C3).

Plain Python / NumPy: nothing here touches the GPU."""
from __future__ import annotations

import csv
import io

import numpy as np


def getrank(v, flattenzeros: bool = True) -> np.ndarray:
    """invperm(sortperm(v)) with 1-based ranks; ties keep their input order (sortperm is stable); with
    `flattenzeros` every zero gets rank 1 (src/utils.jl:16-23)."""
    v = np.asarray(v, dtype=np.float64)
    order = np.argsort(v, kind="stable")
    r = np.empty(v.size, dtype=np.int64)
    r[order] = np.arange(1, v.size + 1)
    if flattenzeros:
        r[v == 0.0] = 1
    return r


def invnormaltransform(v, mu: float = 0.0, sigma: float = 1.0, c: float = 3.0 / 8.0, flattenzeros: bool = True) -> np.ndarray:
    """norminvcdf(mu, sigma, (rank - c) / (n - 2c + 1)) per element (src/utils.jl:25-28)."""
    from scipy.special import ndtri

    r = getrank(v, flattenzeros)
    n = r.size
    return mu + sigma * ndtri((r - c) / (n - 2.0 * c + 1.0))


def read_table(source, delimiter: str | None = None) -> dict[str, list[str]]:
    """Delimited text with a header row -> {column: list of strings}.  `source`: path or text; the delimiter defaults to
    ',' for .csv paths and tab otherwise (the reference's _df_output / CSV.read convention, src/utils.jl:30-40)."""
    if isinstance(source, str) and "\n" not in source:
        if delimiter is None:
            delimiter = "," if source.lower().endswith(".csv") else "\t"
        with open(source, newline="") as f:
            text = f.read()
    else:
        text = source
        if delimiter is None:
            delimiter = "\t" if "\t" in text.splitlines()[0] else ","
    rows = list(csv.reader(io.StringIO(text), delimiter=delimiter))
    if not rows:
        raise ValueError("empty table")
    header, body = rows[0], [r for r in rows[1:] if r]
    for k, r in enumerate(body):
        if len(r) != len(header):
            raise ValueError(f"row {k + 2} has {len(r)} fields, the header has {len(header)}")
    return {h: [r[i] for r in body] for i, h in enumerate(header)}


_MISSING = {"", "NA", "NaN", "nan", "missing", "NULL"}


def complete_cases(table: dict[str, list[str]], columns=None) -> dict[str, list[str]]:
    """Rows without a missing entry in `columns` (default: all) - `df[completecases(df), :]` (CLI/src/select.jl:39)."""
    cols = list(table) if columns is None else list(columns)
    n = len(next(iter(table.values()))) if table else 0
    keep = [i for i in range(n) if all(table[c][i].strip() not in _MISSING for c in cols)]
    return {h: [col[i] for i in keep] for h, col in table.items()}


def column_values(col: list[str]) -> tuple[np.ndarray, list[str] | None]:
    """A numeric column as floats, or a categorical column as dense ids 1..k in order of first appearance (the Cat kernel
    only tests equality, src/gp_parts.jl:11-13) together with its levels."""
    try:
        return np.array([float(x) for x in col], dtype=np.float64), None
    except ValueError:
        levels: dict[str, int] = {}
        ids = np.array([levels.setdefault(x, len(levels) + 1) for x in col], dtype=np.float64)
        return ids, list(levels)


def design_matrix(table: dict[str, list[str]], variables) -> tuple[np.ndarray, dict[str, list[str]]]:
    """`Matrix(df[!, vars])`: one column per entry of `variables` (a variable used by two leaves appears twice, exactly as
    src/abstractgp_translations.jl:45-71 binds leaf i to column i).  Returns (X n x d float64, {categorical var: levels})."""
    cols, levels = [], {}
    for v in variables:
        if v not in table:
            raise KeyError(f"variable {v!r} is not a column of the table (columns: {', '.join(table)})")
        x, lv = column_values(table[v])
        if lv is not None:
            levels[v] = lv
        cols.append(x)
    X = np.column_stack(cols) if cols else np.zeros((0, 0))
    return X, levels


def pack_responses(table: dict[str, list[str]], features, transform: bool = True) -> np.ndarray:
    """The y-batch of config C3: one row per feature column (B x n), optionally inverse-normal transformed per feature."""
    Y = np.array([[float(x) for x in table[f]] for f in features], dtype=np.float64)
    if transform:
        Y = np.vstack([invnormaltransform(row) for row in Y])
    return Y
