"""TEST INFRASTRUCTURE -- CPU restatement of the renderers.

Follows nerfstudio/model_components/renderers.py: RGBRenderer :53-118, AccumulationRenderer :171-197,
DepthRenderer :215-261, SemanticRenderer :284-295 (dense branch only; the packed nerfacc branch is out of scope).
Inputs are plain tensors: rgb [R,S,3], weights [R,S,1], starts/ends [R,S,1].
"""
from typing import Optional, Union

import torch


def render_rgb(rgb, weights, background: Union[str, torch.Tensor], training: bool = False, rand_bg: Optional[torch.Tensor] = None):
    """renderers.py:53-118.  background: tensor[3] | 'last_sample' | 'random' (then ``rand_bg`` [R,3] supplies the draw)."""
    comp = torch.sum(weights * rgb, dim=-2)
    acc = torch.sum(weights, dim=-2)
    if isinstance(background, str):
        if background == "last_sample":
            bg = rgb[..., -1, :]
        elif background == "random":
            bg = rand_bg
        else:
            raise ValueError(background)
    else:
        bg = background
    comp = comp + bg * (1.0 - acc)
    if not training:
        comp = torch.clamp(comp, min=0.0, max=1.0)
    return comp


def render_accumulation(weights):
    """renderers.py:171-197."""
    return torch.sum(weights, dim=-2)


def render_depth(weights, starts, ends, method: str = "expected"):
    """renderers.py:215-261.  NOTE the *batch-global* clip to [steps.min(), steps.max()] in 'expected' (:257)."""
    steps = (starts + ends) / 2
    if method == "median":
        cum = torch.cumsum(weights[..., 0], dim=-1)
        split = torch.ones((*weights.shape[:-2], 1), dtype=weights.dtype) * 0.5
        idx = torch.searchsorted(cum, split, side="left")
        idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
        return torch.gather(steps[..., 0], dim=-1, index=idx)
    if method == "expected":
        eps = 1e-10
        depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
        return torch.clip(depth, steps.min(), steps.max())
    raise NotImplementedError(method)


def render_semantics(semantics, weights):
    """renderers.py:284-295 (used as the normal renderer, base_surface_model.py:216)."""
    return torch.sum(weights * semantics, dim=-2)
