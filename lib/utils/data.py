"""Drop-in for the reference's lib/utils/data.py:3-16 (host->device move of a batch dict)."""
import torch


def data_to_model_device(data, model):
    device = next(model.parameters()).device
    for k, v in data.items():
        if torch.is_tensor(v):
            data[k] = v.to(device)
    return data
