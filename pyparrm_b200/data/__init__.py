"""Example recordings shipped with the package (reference ``src/pyparrm/data``)."""

from .example_data import DATASETS, get_example_data_paths

__all__ = ["DATASETS", "get_example_data_paths"]
