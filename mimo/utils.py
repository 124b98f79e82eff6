from pathlib import Path


def dir_path(string) -> Path:
    """argparse helper: the argument must name an existing directory."""
    p = Path(string)
    if not p.is_dir():
        raise NotADirectoryError(string)
    return p


def count_trainable_parameters(model) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
