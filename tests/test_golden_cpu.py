"""The oracle must keep reproducing the committed golden vectors bit for bit (tests/golden/make_golden.py)."""
import json

import numpy as np

from conftest import GOLDEN, SCENES


def test_oracle_reproduces_golden_vectors(O):
    meta = json.loads((GOLDEN / "golden_v1.json").read_text())
    data = np.load(GOLDEN / "golden_v1.npz")
    assert len(meta["cases"]) == len(data.files) == 8
    for case in meta["cases"]:
        osc = O.OracleScene(str(SCENES / case["object"]), str(SCENES / case["deformation"]) if case["deformation"] else None,
                            flat_field=case["flat_field"], density_multiplier=case["density_multiplier"])
        n = 0
        for v, (az, pol) in enumerate(case["views"]):
            eye, cm = O.camera_from_angles(az, pol, case["R"])
            img, k = osc.render_view(eye, cm, case["res"], case["fov"], case["R"], case["ds"], case["integration"])
            n += k
            assert np.array_equal(img, data[case["key"]][v]), case["key"]
        assert n == case["ref_samples"]
        g = data[case["key"]]
        assert g.min() >= 0.0 and g.max() <= 1.0 and g.min() < 0.95  # a real image, not a blank


def test_scene_fixtures_carry_the_reference_example_values():
    """tests/scenes/*.json are tools/make_scene_fixtures.py re-serialisations of the reference examples."""
    cube = json.loads((SCENES / "cube_w_hole.json").read_text())
    assert [o["type"] for o in cube["objects"]] == ["cube", "sphere", "cylinder"]
    assert cube["objects"][0]["side"] == 1.5 and cube["objects"][1]["rho"] == -1.0
    lat = json.loads((SCENES / "lattice.json").read_text())
    assert len(lat["uc"]["objects"]["objects"]) == 36 and lat["xmax"] == 0.81 and lat["uc"]["xmax"] == 0.4
    gy = json.loads((SCENES / "gyroid_example.json").read_text())
    assert gy["uc"]["objects"]["objects"][0] == {"center": [0.0, 0.0, 0.0], "rho": 1.0, "scale": 0.1, "thickness": 0.2,
                                                 "type": "gyroid"}
    sig = json.loads((SCENES / "deformation_sigmoid.json").read_text())
    assert sig == {"amplitude": 0.2, "center": 0.0, "direction": "z", "lengthscale": 0.2, "type": "sigmoid"}
