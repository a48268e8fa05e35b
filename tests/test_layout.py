import json
import os

from conftest import GOLDEN
from psalm_b200.layout import PhiConfig, PsalmConfig, checkpoint_layout


def test_layout_matches_reference_manifest():
    """Key names / shapes / dtypes equal those of the reference constructors (2-layer Phi build)."""
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifest_phi2layers.json")))
    lay = checkpoint_layout(PsalmConfig(phi=PhiConfig(layers=2)))
    assert set(lay) == set(man)
    for k, (shape, dt, _) in lay.items():
        assert list(shape) == man[k][0] and dt == man[k][1], k


def test_full_size_parameter_count():
    import math
    n = sum(math.prod(s) for s, _, _ in checkpoint_layout().values())
    assert abs(n / 1e6 - 1591.27) < 0.5  # 1590.8 M (+ buffers)
