import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    import snch_lbvh_b200 as p
    return p


@pytest.fixture(scope="session")
def meshes(pkg):
    return pkg.meshes


def small_cases(m):
    """name -> (verts, tris): the meshes every parity test walks (closed, open/boundary, collisions, tiny)."""
    return {
        "tet": m.tetrahedron(),
        "ico2": m.icosphere(2),
        "grid6": m.open_grid(6),
        "torus24x16": m.bumpy_torus(24, 16),
    }


def small_cases2(m):
    """name -> (verts, segments): 2-D polylines (closed, open with free ends, shuffled / inconsistently oriented soup).
    A one-segment scene is not here: the reference's traversals read out of bounds on it."""
    return {
        "poly_circle": m.wavy_circle(257, 5, 0.2),
        "poly_open": m.open_polyline(97),
        "poly_soup": m.polyline_soup(3, 4, 60),
    }
