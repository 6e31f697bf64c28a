"""Streaming data access (mirror of reference ``torch_rechub/data``): Parquet files -> dicts of column tensors."""
from .convert import pa_array_to_tensor
from .dataset import ParquetIterableDataset

__all__ = ["ParquetIterableDataset", "pa_array_to_tensor"]
