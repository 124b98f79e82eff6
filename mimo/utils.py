"""Small host-side helpers of the reference's `mimo.utils` surface (reference: mimo/utils.py): an argparse type for
directories and the trainable-parameter count that MimoUnetModel stores in its hyper-parameters."""
import os
import pathlib


def dir_path(string) -> pathlib.Path:
    """argparse `type=`: accepts only names of existing directories, returns them as a Path."""
    if os.path.isdir(string):
        return pathlib.Path(string)
    raise NotADirectoryError(string)


def count_trainable_parameters(model) -> int:
    """Number of scalar parameters that receive gradients."""
    total = 0
    for param in model.parameters():
        if param.requires_grad:
            total += param.numel()
    return total
