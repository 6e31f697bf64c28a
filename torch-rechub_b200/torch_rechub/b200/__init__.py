"""B200 (sm_100a) engine under torch-rechub's Python API.

``torch_rechub.b200`` holds everything that is not part of the reference's public surface:
the ctypes binding of the C-ABI library (``_lib``), the table gradient manager (``table``), the
autograd wrappers around the kernels (``ops``), the row-wise optimisers (``optim``), the CUDA-graph
step runner (``graph``) and the multi-GPU field sharding (``dist``).

CUDA tensors ALWAYS go through ``librechub_b200.so``; when the library is missing the first CUDA
call raises (there is no silent eager fallback).  CPU tensors follow the reference's own
composite-of-torch-ops arithmetic (the quick-start / ONNX-export configuration of the reference).
"""
from . import config  # noqa: F401
