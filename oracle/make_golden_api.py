"""TEST INFRASTRUCTURE -- mints tests/golden/reference_api.json from the UNMODIFIED reference: SDFFieldConfig field defaults
(nerfstudio/fields/sdf_field.py:121-185) and the constructor signatures (parameter names + defaults) of the samplers, renderers,
colliders and density field the drop-in mirrors.  tests/test_abi_cpu.py compares the product's classes against it.

    python -m oracle.make_golden_api
"""
import dataclasses
import inspect
import json
import os

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_api.json")


def _plain(v):
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    if isinstance(v, (tuple, list)):
        return [_plain(x) for x in v]
    return f"<{type(v).__name__}>"


def signature(cls):
    out = []
    for name, p in inspect.signature(cls.__init__).parameters.items():
        if name in ("self", "kwargs", "args"):
            continue
        out.append([name, None if p.default is inspect.Parameter.empty else _plain(p.default)])
    return out


def main():
    ref_import.install_shims()
    import warnings

    warnings.simplefilter("ignore")
    from nerfstudio.fields import density_fields, sdf_field
    from nerfstudio.model_components import ray_samplers, renderers, scene_colliders

    api = {"SDFFieldConfig": {f.name: _plain(f.default) for f in dataclasses.fields(sdf_field.SDFFieldConfig)
                              if f.name != "_target" and f.default is not dataclasses.MISSING}}
    for mod, names in ((ray_samplers, ["SpacedSampler", "UniformSampler", "LinearDisparitySampler", "SqrtSampler", "LogSampler",
                                       "UniformLinDispPiecewiseSampler", "PDFSampler", "ProposalNetworkSampler", "ErrorBoundedSampler", "NeuSSampler",
                                       "UniSurfSampler"]),
                       (renderers, ["RGBRenderer", "DepthRenderer"]),
                       (scene_colliders, ["AABBBoxCollider", "NearFarCollider", "SphereCollider"]),
                       (density_fields, ["HashMLPDensityField"]),
                       (sdf_field, ["SDFField", "LaplaceDensity", "SingleVarianceNetwork"])):
        for n in names:
            api[n] = signature(getattr(mod, n))
    with open(OUT, "w") as fh:
        json.dump(api, fh, indent=1, sort_keys=True)
    print("wrote", OUT, sorted(api))


if __name__ == "__main__":
    main()
