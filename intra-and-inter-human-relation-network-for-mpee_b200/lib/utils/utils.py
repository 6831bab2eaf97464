"""`from utils.utils import get_valid_output` -- the one helper of lib/utils on the forward path."""
import torch


def get_valid_output(outputs, length):
    """Drop zero-padded persons: [bs*max(length), ...] -> [sum(length), ...]  (reference: lib/utils/utils.py:24-37).

    The B200 modules keep ragged (un-padded) batches internally, so they never call this; it is
    provided for callers that build padded tensors themselves.
    """
    n_max = max(length)
    rest = tuple(outputs.shape[1:])
    grouped = outputs.reshape((outputs.shape[0] // n_max, n_max) + rest)
    return torch.cat([grouped[i, :n] for i, n in enumerate(length)], dim=0)
