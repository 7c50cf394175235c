"""Drop-in for the reference's `lib/utils/data.py` (`data_to_model_device`, :3-16): the batch dict's tensors follow the
model to its device; everything else (names, ids, lists) stays as it is.  A model without parameters keeps the batch on
the CPU, like the reference's fallback for its parameter-free baselines."""
import torch


def _device_of(model) -> torch.device:
    first = next(iter(model.parameters()), None) if hasattr(model, "parameters") else None
    return first.device if first is not None else torch.device("cpu")


def data_to_model_device(data, model):
    target = _device_of(model)
    data.update({key: val.to(target) for key, val in data.items() if isinstance(val, torch.Tensor)})
    return data
