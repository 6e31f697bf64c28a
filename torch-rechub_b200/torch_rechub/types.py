"""Shared type aliases (mirror of reference ``torch_rechub/types.py``)."""
import os
import typing as ty

#: Path to a file.
FilePath = ty.Union[str, os.PathLike]
