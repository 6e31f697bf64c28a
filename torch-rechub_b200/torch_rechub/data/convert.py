"""PyArrow arrays -> torch tensors (mirror of reference ``torch_rechub/data/convert.py:10-67``).

Same contract as the reference: booleans / integers / floats / nulls become a 1-D ``float32`` tensor; list columns (list, large_list,
fixed_size_list) of those become ``(rows, width)`` ``float32`` and must be rectangular.  Ids therefore arrive as floats — the models
truncate them with ``.long()`` exactly as the reference does (basic/layers.py:83); the engine's gather takes that cast."""
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.types as pt
import torch


def _scalar_ok(t):
    return pt.is_boolean(t) or pt.is_integer(t) or pt.is_floating(t) or pt.is_null(t)


def _list_ok(t):
    return pt.is_list(t) or pt.is_large_list(t) or pt.is_fixed_size_list(t)


def _numpy(arr):
    # a writable copy: torch.from_numpy refuses PyArrow's read-only zero-copy views
    return arr.to_numpy(zero_copy_only=False, writable=True)


def pa_array_to_tensor(arr):
    """``pa.Array`` -> ``torch.Tensor`` (float32).  ``TypeError`` for unsupported (value) types, ``ValueError`` for ragged lists."""
    t = arr.type
    if _scalar_ok(t):
        return torch.from_numpy(_numpy(pc.cast(arr, pa.float32())))
    if not _list_ok(t):
        raise TypeError(f"Unsupported array type: {t}")
    if not _scalar_ok(t.value_type):
        raise TypeError(f"Unsupported value type in the nested array: {t.value_type}")
    if len(pc.unique(pc.list_value_length(arr))) > 1:
        raise ValueError("Cannot convert the ragged nested array.")
    flat = _numpy(pc.cast(arr, pa.list_(pa.float32())).flatten())  # flatten() honours a sliced batch's offsets (``.values`` does not)
    rows = len(arr)
    return torch.from_numpy(flat.reshape(rows, -1 if rows > 0 else 0))  # an empty list-of-lists is (0, 0)
