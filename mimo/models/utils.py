"""MIMO input transforms and ensemble aggregation with the reference's interface
(reference: mimo/models/utils.py)."""
from typing import List, Optional

import torch


def shuffle_indices(batch: int, num_subnetworks: int, input_repetition_probability: float = 0.0,
                    batch_repetitions: int = 1, device=None) -> List[torch.Tensor]:
    """Per-subnetwork batch permutations of apply_input_transform (reference utils.py:27-36): one main
    permutation (repeated `batch_repetitions` times); its first (1 - p_rep) share is re-shuffled independently
    for every subnetwork with the CPU generator, the tail stays aligned across subnetworks."""
    main = torch.randperm(batch, device=device).repeat(batch_repetitions)
    k = int(main.shape[0] * (1.0 - input_repetition_probability))
    out = []
    for _ in range(num_subnetworks):
        perm = torch.randperm(k)  # CPU generator, like the reference
        if main.is_cuda:
            # indexing a CUDA tensor with a pageable CPU index makes PyTorch copy it with a blocking cudaStreamSynchronize
            # (twice per training step: the GPU drains while the host enqueues the next forward). Pinned + non_blocking
            # keeps the step asynchronous; the values are identical.
            perm = perm.pin_memory().to(main.device, non_blocking=True)
        out.append(torch.cat((main[:k][perm], main[k:]), dim=0))
    return out


def apply_input_transform(image: torch.Tensor, label: torch.Tensor, mask: Optional[torch.Tensor], num_subnetworks: int,
                          input_repetition_probability: float = 0.0, batch_repetitions: int = 1):
    """[B,C,H,W] -> [B*rep, S, C, H, W] for image, label and (optional) mask."""
    idx = shuffle_indices(image.shape[0], num_subnetworks, input_repetition_probability, batch_repetitions, image.device)

    def gather(t):
        return None if t is None else torch.stack([t.index_select(0, i) for i in idx], dim=1)

    return gather(image), gather(label), gather(mask)


def repeat_subnetworks(x: torch.Tensor, num_subnetworks: int):
    """[B,C,H,W] -> [B,S,C,H,W] (materialised, like the reference)."""
    return x[:, None].repeat(1, num_subnetworks, 1, 1, 1)


def flatten_subnetwork_dimension(x: torch.Tensor):
    """[B,S,C,H,W] -> [B*S,C,H,W]."""
    b, s, c, h, w = x.shape
    return x.reshape(b * s, c, h, w)


def compute_uncertainties(criterion, y_preds, log_params):
    """mean, aleatoric variance, epistemic variance over the subnetwork dimension, each [B,C,H,W].

    Laplace members on the GPU use the fused aggregation kernel (mean, mean 2 b^2, unbiased variance);
    anything else follows the reference formula with torch ops."""
    from mimo.losses import LaplaceNLL
    if isinstance(criterion, LaplaceNLL) and y_preds.is_cuda:
        from mimo_unet_b200 import functional as Fn
        return Fn.ensemble_aggregate(y_preds, log_params)
    S = y_preds.shape[1]
    mean = criterion.mode(y_preds, log_params).mean(dim=1)
    aleatoric = torch.square(criterion.std(y_preds, log_params)).mean(dim=1)
    if S > 1:
        epistemic = torch.square(y_preds - y_preds.mean(dim=1, keepdim=True)).sum(dim=1) / (S - 1)
    else:
        epistemic = torch.zeros_like(aleatoric)
    return mean, aleatoric, epistemic
