"""speechbrain.nnet.losses subset [SB-recall, 0.5.x].

length_to_mask is the only function whose arithmetic the reference's hot path
uses (data_utils.py:88); compute_masked_loss is imported by vanilla_vae.py:4
and models/test_vanilla_vae/model.py:3 but never called."""
import torch


def length_to_mask(length, max_len=None, dtype=None, device=None):
    assert len(length.shape) == 1
    if max_len is None:
        max_len = length.max().long().item()
    mask = torch.arange(max_len, device=length.device, dtype=length.dtype).expand(
        len(length), max_len) < length.unsqueeze(1)
    if dtype is None:
        dtype = length.dtype
    if device is None:
        device = length.device
    return torch.as_tensor(mask, dtype=dtype, device=device)


def compute_masked_loss(*args, **kwargs):
    raise NotImplementedError("shim: never called on the hot path")
