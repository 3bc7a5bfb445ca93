"""TEST INFRASTRUCTURE -- mints tests/golden/sdffield_state_keys.json: the state_dict key -> shape map of the UNMODIFIED reference
SDFField (nerfstudio/fields/sdf_field.py) for every named case, so that checkpoint compatibility of the drop-in (same names, same
shapes) is pinned without needing the reference at test time.

    python -m oracle.make_golden_statedict
"""
import json
import os

from . import cases, ref_import
from .field import init_params
from .make_golden import build_reference_field

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sdffield_state_keys.json")


def main():
    R = ref_import.ref_modules()
    out = {}
    for name in cases.CASES:
        spec, kw = cases.CASES[name]
        params = init_params(spec, **cases.init_kwargs(kw))
        f = build_reference_field(R, spec, params, kw)
        # the hash grid lives behind the tinycudann boundary: with real tcnn the key is `encoding.params` (flat); the stand-in's
        # `encoding.enc.hash_table` is an artefact of this container and is left out
        out[name] = {k: list(v.shape) for k, v in f.state_dict().items() if not k.startswith("encoding.")}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
