"""Loading existing sdfstudio checkpoints into the drop-in modules.

A reference ``step-*.ckpt`` (engine/trainer.py:276-297) is ``{"step", "pipeline", "optimizers", "schedulers", "scalers"}`` where
``pipeline`` is ``Pipeline.state_dict()``: the field's tensors sit under ``_model.field.`` (``module.`` in front when the pipeline
was DDP-wrapped, pipelines/base_pipeline.py:426-439).  SDFField keeps the reference's parameter names and shapes
(tests/test_abi_cpu.py::test_state_dict_names_match_reference), so loading is a prefix strip plus the hash-grid entry:

* ``encoding.params`` -- tiny-cuda-nn's single flat parameter vector (fp32 master copy; older builds store fp16): loaded into
  ``Encoding(layout="tcnn").params`` (same level order / per-level sizes, encoding.py ``make_grid_desc``), cast to fp32.
* ``encoding.hash_table`` -- the reference's own torch ``HashEncoding`` ([L*T, F]): loaded into ``layout="torch"``.
"""
from typing import Dict, Tuple

import torch

FIELD_PREFIX = "_model.field."


def extract_state(loaded_state: Dict, prefix: str = FIELD_PREFIX) -> Dict[str, torch.Tensor]:
    """checkpoint dict (or its ``pipeline`` entry, or an already flat state_dict) -> tensors under `prefix`, prefix removed."""
    state = loaded_state.get("pipeline", loaded_state) if isinstance(loaded_state, dict) else loaded_state
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}
    sub = {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
    return sub if sub else dict(state)


def load_field_checkpoint(field, loaded_state, prefix: str = FIELD_PREFIX, strict: bool = True) -> Tuple[list, list]:
    """Load a reference checkpoint (path, checkpoint dict or state_dict) into a ``sdfstudio_b200.SDFField``.
    Returns (missing, unexpected) like ``load_state_dict``; with ``strict`` any mismatch other than the read-only ``aabb`` raises."""
    if isinstance(loaded_state, (str, bytes)) or hasattr(loaded_state, "__fspath__"):
        loaded_state = torch.load(loaded_state, map_location="cpu")
    sd = extract_state(loaded_state, prefix)
    enc = field.encoding
    if "encoding.params" in sd:
        if enc.layout != "tcnn":
            raise ValueError("the checkpoint holds a tiny-cuda-nn grid (`encoding.params`); build the field with grid_layout='tcnn'")
        flat = sd["encoding.params"].reshape(-1).to(torch.float32)
        if flat.numel() != enc.params.numel():
            raise ValueError(f"encoding.params has {flat.numel()} entries, this grid configuration needs {enc.params.numel()} "
                             "(check num_levels / log2_hashmap_size / base_res / max_res / hash_features_per_level)")
        sd["encoding.params"] = flat
    if "encoding.hash_table" in sd and enc.layout != "torch":
        raise ValueError("the checkpoint holds a torch HashEncoding table (`encoding.hash_table`); build the field with grid_layout='torch'")
    sd = {k: (v.to(torch.float32) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sd.items()}
    missing, unexpected = field.load_state_dict(sd, strict=False)
    missing = [m for m in missing if m != "aabb"]
    if strict and (missing or unexpected):
        raise RuntimeError(f"checkpoint does not match the field: missing {missing}, unexpected {list(unexpected)}")
    return missing, list(unexpected)


def load_density_field_checkpoint(density_field, loaded_state, index: int = 0, strict: bool = True):
    """Load ``_model.proposal_networks.{index}.mlp_base.params`` (tiny-cuda-nn ``NetworkWithInputEncoding``: FullyFusedMLP weights then the
    HashGrid table, nerfstudio/fields/density_fields.py:89-96) of a reference neus-facto / bakedsdf checkpoint into a
    ``sdfstudio_b200.HashMLPDensityField``.  The vector length is checked (the output layer is stored 16 rows wide like tcnn pads it); the
    ordering inside the vector follows tcnn's published layout (UNPINNED: tcnn is not vendored in the reference)."""
    if isinstance(loaded_state, (str, bytes)) or hasattr(loaded_state, "__fspath__"):
        loaded_state = torch.load(loaded_state, map_location="cpu")
    sd = extract_state(loaded_state, f"_model.proposal_networks.{index}.")
    key = "mlp_base.params"
    if key not in sd:
        raise KeyError(f"{key} not found under _model.proposal_networks.{index}.")
    flat = sd[key].reshape(-1).to(torch.float32)
    nb = density_field.mlp_base
    if flat.numel() != nb.params.numel():
        raise ValueError(f"mlp_base.params has {flat.numel()} entries, this proposal network needs {nb.params.numel()} "
                         "(check num_levels / log2_hashmap_size / max_res / hidden_dim / num_layers)")
    with torch.no_grad():
        nb.params.copy_(flat.to(nb.params.device))
    extra = [k for k in sd if k not in (key, "aabb")]
    if strict and extra:
        raise RuntimeError(f"unexpected entries for the proposal network: {extra}")
    return [], extra
