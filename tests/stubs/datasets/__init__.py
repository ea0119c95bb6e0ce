"""TEST INFRASTRUCTURE: JSON-backed stand-in for the few `datasets` calls the reference's embedding scripts make
(see ../README.md).  File access goes through builtins.open / os so that the `modal` shim's volume mounts apply."""
import json
import os

__version__ = "0.0-stub"


class Dataset:
    def __init__(self, table_or_rows):
        self._table = None
        if isinstance(table_or_rows, list):
            self._rows = table_or_rows
        else:  # a pyarrow.Table (wikipedia/main.py:203-218)
            self._table = table_or_rows
            self._rows = None

    def __len__(self):
        return len(self._rows) if self._rows is not None else self._table.num_rows

    def __iter__(self):
        return iter(self._rows if self._rows is not None else self._table.to_pylist())

    def __getitem__(self, k):
        if isinstance(k, str):
            return [r[k] for r in self]
        return (self._rows if self._rows is not None else self._table.to_pylist())[k]

    def select(self, indices):
        rows = self._rows if self._rows is not None else self._table.to_pylist()
        return Dataset([rows[i] for i in indices])

    def save_to_disk(self, path):
        import pyarrow as pa
        import pyarrow.parquet as pq

        table = self._table if self._table is not None else pa.Table.from_pylist(self._rows)
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "data.parquet"), "wb") as f:
            pq.write_table(table, f)
        with open(os.path.join(path, "dataset_info.json"), "w") as f:
            json.dump({"num_rows": table.num_rows, "columns": table.column_names}, f)


class DatasetDict(dict):
    pass


def _read(path):
    with open(path) as f:
        return Dataset(json.load(f))


def load_from_disk(path):
    out = DatasetDict()
    for name in sorted(os.listdir(path)):
        if name.endswith(".json"):
            out[name[:-5]] = _read(os.path.join(path, name))
    if not out:
        raise FileNotFoundError(f"no <split>.json under {path}")
    return out


def load_dataset(name, subset=None, split=None, **_kw):
    root = os.environ["FAKE_DATASETS_DIR"]
    return _read(os.path.join(root, f"{subset or name.replace('/', '--')}.{split or 'train'}.json"))
