"""FieldHeadNames (nerfstudio/field_components/field_heads.py:28-45).  Inside sdfstudio the reference enum is used so that
dictionary keys compare equal; stand-alone an identical enum is defined."""
try:  # pragma: no cover - only when running inside the reference code base
    from nerfstudio.field_components.field_heads import FieldHeadNames  # type: ignore
except Exception:  # noqa: BLE001
    from enum import Enum

    class FieldHeadNames(Enum):
        RGB = "rgb"
        SH = "sh"
        DENSITY = "density"
        NORMALS = "normals"
        PRED_NORMALS = "pred_normals"
        UNCERTAINTY = "uncertainty"
        TRANSIENT_RGB = "transient_rgb"
        TRANSIENT_DENSITY = "transient_density"
        SEMANTICS = "semantics"
        NORMAL = "normal"
        SDF = "sdf"
        ALPHA = "alpha"
        GRADIENT = "gradient"
        OCCUPANCY = "occupancy"
