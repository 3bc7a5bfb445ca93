"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (sdfstudio @ /root/reference) on CPU.

Only usable inside the build container (``/root/reference`` does not exist on the GPU box).  It is used by
``oracle/make_golden.py`` to mint the golden vectors under ``tests/golden/`` and by the container-only tests that
pin ``oracle/*.py`` (the CPU restatement) against the real reference code.  Nothing in the product path
(``sdfstudio_b200/``) may import this module.

Shims (nothing here re-implements reference arithmetic except the ``tinycudann`` stand-in, which is assembled from the
reference's own classes):

* ``torchtyping``  -- annotation-only package, absent here.
* ``nerfacc``      -- imported at module top of ray_samplers.py:23-25 / renderers.py:32, unused by in-scope classes.
* ``nerfstudio.configs.base_config`` -- the real file fails to import on py>=3.11 (base_config.py:125); only
  ``InstantiateConfig`` / ``PrintableConfig`` are needed (fields/base_field.py:27).
* ``tinycudann``   -- ``SDFField`` hard-requires ``tcnn.Encoding`` (sdf_field.py:228-241).  The stand-in is backed by the
  reference's own pure-PyTorch ``HashEncoding.pytorch_fwd`` (encodings.py:357-398) plus the smoothstep line of
  ``PeriodicVolumeEncoding.pytorch_fwd`` (encodings.py:700-701).  This is BASELINE.json config[0]'s
  "pure-PyTorch HashEncoding (use_tcnn=False) on CPU".
"""
import dataclasses
import os
import sys
import types
from typing import Any, Type

REFERENCE_ROOT = os.environ.get("SDFSTUDIO_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nerfstudio"))


_done = False


def install_shims():
    """Make ``import nerfstudio.fields.sdf_field`` & friends work on this CPU-only py3.12 container."""
    global _done
    if _done:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    from torch import nn

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    class _TT:  # torchtyping.TensorType[...] -> itself
        def __class_getitem__(cls, item):
            return cls

    m = types.ModuleType("torchtyping")
    m.TensorType = _TT
    sys.modules["torchtyping"] = m

    m = types.ModuleType("nerfacc")
    m.OccupancyGrid = object
    sys.modules["nerfacc"] = m

    bc = types.ModuleType("nerfstudio.configs.base_config")

    class PrintableConfig:  # pylint: disable=too-few-public-methods
        pass

    @dataclasses.dataclass
    class InstantiateConfig(PrintableConfig):
        _target: Type

        def setup(self, **kw) -> Any:
            return self._target(self, **kw)

    bc.PrintableConfig, bc.InstantiateConfig = PrintableConfig, InstantiateConfig
    import nerfstudio.configs  # noqa: F401  (package __init__ is empty)

    sys.modules["nerfstudio.configs.base_config"] = bc

    # ---- tinycudann stand-in built from the reference's own torch HashEncoding --------------------------------
    from nerfstudio.field_components.encodings import HashEncoding

    class _RefTorchHashGrid(nn.Module):
        """tcnn.Encoding(n_input_dims=3, encoding_config=...) look-alike (sdf_field.py:230-241 call site)."""

        def __init__(self, n_input_dims, encoding_config, seed=None, dtype=None):
            super().__init__()
            assert n_input_dims == 3
            cfg = encoding_config
            L = int(cfg["n_levels"])
            base = int(cfg["base_resolution"])
            g = float(cfg["per_level_scale"])
            max_res = base * g ** (L - 1)
            # HashEncoding recomputes growth_factor from (min_res, max_res); feed it max_res so that it recovers g.
            self.enc = HashEncoding(
                num_levels=L,
                min_res=base,
                max_res=max_res,
                log2_hashmap_size=int(cfg["log2_hashmap_size"]),
                features_per_level=int(cfg["n_features_per_level"]),
                implementation="torch",
            )
            self.smoothstep = cfg.get("interpolation", "Linear") == "Smoothstep"
            self.n_output_dims = self.enc.get_out_dim()

        def forward(self, x):
            if not self.smoothstep:
                return self.enc.pytorch_fwd(x)
            return _smooth_fwd(self.enc, x)

    def _smooth_fwd(enc, in_tensor):
        # HashEncoding.pytorch_fwd (encodings.py:357-398) with the smoothstep remap of
        # PeriodicVolumeEncoding.pytorch_fwd (encodings.py:700-701) applied to `offset`.  The corner/hash/blend
        # expressions are executed by re-using the reference methods; only `offset` is altered.
        in_tensor = in_tensor[..., None, :]
        scaled = in_tensor * enc.scalings.view(-1, 1).to(in_tensor.device)
        scaled_c = torch.ceil(scaled).type(torch.int32)
        scaled_f = torch.floor(scaled).type(torch.int32)
        offset = scaled - scaled_f
        offset = offset * offset * (3.0 - 2.0 * offset)
        c, f = scaled_c, scaled_f
        cat = torch.cat
        h = enc.hash_fn
        hashed = [
            h(c),
            h(cat([c[..., 0:1], f[..., 1:2], c[..., 2:3]], dim=-1)),
            h(cat([f[..., 0:1], f[..., 1:2], c[..., 2:3]], dim=-1)),
            h(cat([f[..., 0:1], c[..., 1:2], c[..., 2:3]], dim=-1)),
            h(cat([c[..., 0:1], c[..., 1:2], f[..., 2:3]], dim=-1)),
            h(cat([c[..., 0:1], f[..., 1:2], f[..., 2:3]], dim=-1)),
            h(f),
            h(cat([f[..., 0:1], c[..., 1:2], f[..., 2:3]], dim=-1)),
        ]
        f_0, f_1, f_2, f_3, f_4, f_5, f_6, f_7 = (enc.hash_table[i] for i in hashed)
        f_03 = f_0 * offset[..., 0:1] + f_3 * (1 - offset[..., 0:1])
        f_12 = f_1 * offset[..., 0:1] + f_2 * (1 - offset[..., 0:1])
        f_56 = f_5 * offset[..., 0:1] + f_6 * (1 - offset[..., 0:1])
        f_47 = f_4 * offset[..., 0:1] + f_7 * (1 - offset[..., 0:1])
        f0312 = f_03 * offset[..., 1:2] + f_12 * (1 - offset[..., 1:2])
        f4756 = f_47 * offset[..., 1:2] + f_56 * (1 - offset[..., 1:2])
        encoded_value = f0312 * offset[..., 2:3] + f4756 * (1 - offset[..., 2:3])
        return torch.flatten(encoded_value, start_dim=-2, end_dim=-1)

    t = types.ModuleType("tinycudann")
    t.Encoding = _RefTorchHashGrid
    sys.modules["tinycudann"] = t
    _done = True


def ref_modules():
    """Returns a namespace with the reference classes on the hot path."""
    install_shims()
    from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples
    from nerfstudio.field_components import encodings, spatial_distortions
    from nerfstudio.field_components.field_heads import FieldHeadNames
    from nerfstudio.fields import sdf_field
    from nerfstudio.model_components import ray_samplers, renderers

    ns = types.SimpleNamespace(
        Frustums=Frustums,
        RayBundle=RayBundle,
        RaySamples=RaySamples,
        encodings=encodings,
        spatial_distortions=spatial_distortions,
        FieldHeadNames=FieldHeadNames,
        sdf_field=sdf_field,
        ray_samplers=ray_samplers,
        renderers=renderers,
    )
    return ns
