"""Registry of the example recordings (reference ``src/pyparrm/data/example_data.py:6-30``).

The ``.npy`` files under ``example_data/`` are byte-identical copies of the reference's data
files (made by ``oracle/make_golden.py``); they are inputs and known answers, not code.
"""

from pathlib import Path

DATASETS = {
    "example_data": "example_data.npy",
    "example_data_artefact_free": "example_data_artefact_free.npy",
    "matlab_filtered": "matlab_filtered.npy",
    "ecog_lfp_data": "ecog_lfp_data.npy",
}

_DATA_DIR = Path(__file__).resolve().parent / "example_data"


def get_example_data_paths(name: str) -> str:
    """Path of the example recording called ``name``."""
    if name not in DATASETS:
        raise ValueError(f"`name` must be one of: {list(DATASETS.keys())}")
    return str(_DATA_DIR / DATASETS[name])
