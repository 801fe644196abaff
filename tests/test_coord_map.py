"""on-disk formats of the registration output and the pairwise
motion-distance map (reference PointCloud/helper_functions.py:10-45, coord_map.py:186-307).

CPU tests pin the numpy oracle against the reference's own output (tests/golden/coord_map.npz,
made by tests/golden/make_golden_coord_map.py) and against scipy's Rotation; GPU tests compare the
CUDA operator (through the C ABI) with the oracle and the golden vectors.  Floating point:
tolerance 1e-9 absolute on distances of order 1 (north_star: 1e-5 on transforms)."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

TOL = 1e-9


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "coord_map.npz"))


def _track(T, K, seed):
    """a pose track like the registration output: smooth per-cluster rigid motion"""
    rng = np.random.default_rng(seed)
    M = np.tile(np.eye(4), (T, K, 1, 1))
    R0 = Rotation.random(K, random_state=seed).as_matrix()
    t0 = rng.uniform(-0.3, 0.3, (K, 3))
    ax = rng.normal(size=(K, 3))
    ax /= np.linalg.norm(ax, axis=1)[:, None]
    rate, vel = np.deg2rad(rng.uniform(0, 9, K)), rng.normal(scale=0.01, size=(K, 3))
    for t in range(T):
        for k in range(K):
            M[t, k, :3, :3] = Rotation.from_rotvec(ax[k] * rate[k] * t).as_matrix() @ R0[k]
            M[t, k, :3, 3] = t0[k] + vel[k] * t
    return M


# ------------------------------------------------------------------ CPU: the oracle and what pins it
def test_oracle_matches_reference_coord_dist_map_golden(gold):
    from oracle import coord_map_oracle as C
    for d in (1, 0):
        m, s = C.coord_dist_map(gold["matrices"], float(gold["bounding_box"]), diff=bool(d))
        assert m.shape == gold[f"map_diff{d}"].shape
        assert np.abs(m - gold[f"map_diff{d}"]).max() <= 1e-13
        assert np.abs(s - gold[f"sum_diff{d}"]).max() <= 1e-13


def test_oracle_roma_restatement_vs_scipy():
    from oracle import coord_map_oracle as C
    a, b = Rotation.random(400, random_state=1), Rotation.random(400, random_state=2)
    assert np.abs(C.rotmat_to_rotvec(a.as_matrix()) - a.as_rotvec()).max() <= 1e-13
    ang = (a.inv() * b).magnitude()
    assert np.abs(C.rotvec_geodesic_distance(a.as_rotvec(), b.as_rotvec()) - ang).max() <= 1e-12
    assert np.abs(C.rotmat_geodesic_distance(a.as_matrix(), b.as_matrix()) - ang).max() <= 1e-11
    tiny = Rotation.from_rotvec(np.array([[1e-5, -2e-5, 3e-6], [0, 0, 0], [np.pi, 0, 0]]))   # series branch, identity, half turn
    assert np.abs(C.rotmat_to_rotvec(tiny.as_matrix()) - tiny.as_rotvec()).max() <= 1e-13


def test_file_formats_round_trip_and_reference_layout(gold, tmp_path):
    from oracle import coord_map_oracle as C
    from autourdf_b200 import helper_functions as H     # numpy-only functions: no device needed
    assert list(gold["npz_keys"]) == [str(i) for i in range(11)]        # stored order, not lexicographic
    assert np.array_equal(gold["npz_sizes_in"], gold["npz_sizes_back"])
    segs = np.split(gold["npz_concat_in"], np.cumsum(gold["npz_sizes_in"])[:-1])
    for save, load in ((H.save_pc_npz, C.load_pc_npz), (C.save_pc_npz, H.load_pc_npz)):
        p = str(tmp_path / "0000.npz")
        save(segs, p)
        back = load(p)
        assert [b.shape for b in back] == [s.shape for s in segs]
        assert np.array_equal(np.concatenate(back), gold["npz_concat_back"])
    (tmp_path / "matrix").mkdir()
    for t in range(gold["matrices"].shape[0]):
        np.save(str(tmp_path / "matrix" / f"{t:04}.npy"), gold["matrices"][t])
    rot, mats = C.load_matrix(str(tmp_path) + "/", 0, 7)
    assert np.abs(rot - gold["load_matrix_rot"]).max() <= 1e-15 and np.array_equal(mats, gold["load_matrix_matrices"])
    rot, _ = C.load_matrix(str(tmp_path) + "/", 2, 5)
    assert np.abs(rot - gold["load_matrix_rot_2_5"]).max() <= 1e-15
    assert C.load_matrix(str(tmp_path) + "/")[1].size == 0            # end_steps=0 selects nothing, as upstream


# ------------------------------------------------------------------ GPU: the CUDA operator
@pytest.mark.gpu
def test_coord_dist_map_golden(gold):
    from autourdf_b200.coord_map import coord_dist_map
    for d in (1, 0):
        m, s = coord_dist_map(gold["matrices"], float(gold["bounding_box"]), diff=bool(d))
        assert m.dtype == np.float64 and m.shape == gold[f"map_diff{d}"].shape
        assert np.abs(m - gold[f"map_diff{d}"]).max() <= TOL
        assert np.abs(s - gold[f"sum_diff{d}"]).max() <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("T,K", [(10, 20), (50, 48), (3, 1), (2, 130), (100, 128)])
def test_coord_dist_map_vs_oracle(T, K):
    from oracle import coord_map_oracle as C
    from autourdf_b200.coord_map import coord_dist_map
    M = _track(T, K, seed=T * 1000 + K)
    for d in (True, False):
        m, s = coord_dist_map(M, 0.9, diff=d)
        mo, so = C.coord_dist_map(M, 0.9, diff=d)
        assert m.shape == mo.shape
        assert np.abs(m - mo).max() <= TOL and np.abs(s - so).max() <= TOL * max(T, 1)
        assert np.abs(m - np.transpose(m, (1, 0, 2))).max() <= 1e-12        # symmetric in (j, k)
        assert np.abs(np.einsum("jji->ji", m)).max() <= 1e-12                 # zero diagonal


@pytest.mark.gpu
def test_coord_dist_map_edge_cases():
    import torch
    from autourdf_b200 import _lib
    from autourdf_b200.coord_map import coord_dist_map
    M = _track(1, 5, seed=3)
    m, s = coord_dist_map(M, 1.0, diff=True)          # one frame: no motion steps
    assert m.shape == (5, 5, 0) and np.array_equal(s, np.zeros((5, 5)))
    m, s = coord_dist_map(torch.as_tensor(_track(4, 6, 1)).cuda(), 1.0)       # tensor in, tensor out
    assert torch.is_tensor(m) and m.is_cuda and m.shape == (6, 6, 3)
    with pytest.raises(_lib.AurdfError):
        coord_dist_map(M, 0.0)
    with pytest.raises(_lib.AurdfError):
        coord_dist_map(_track(2, 400, 1), 1.0)


@pytest.mark.gpu
def test_coordmap_loader_and_pose_vectors(gold, tmp_path):
    import torch
    from autourdf_b200 import helper_functions as H
    from autourdf_b200.coord_map import CoordMap
    (tmp_path / "matrix").mkdir()
    (tmp_path / "cluster").mkdir()
    for t in range(7):
        np.save(str(tmp_path / "matrix" / f"{t:04}.npy"), gold["matrices"][t])
        H.save_pc_npz([np.full((k + 1, 3), float(t)) for k in range(9)], str(tmp_path / "cluster" / f"{t:04}.npz"))
    cm = CoordMap(str(tmp_path) + "/", float(gold["bounding_box"]), 0, 7)
    assert np.abs(cm.coords - gold["load_matrix_rot"]).max() <= 1e-12 and cm.num_coords == 9
    assert len(cm.clusters) == 7 and cm.clusters[3]["8"].shape == (9, 3)
    m, s = cm.coord_dist_map(diff=True)
    assert np.abs(m - gold["map_diff1"]).max() <= TOL and np.abs(s - gold["sum_diff1"]).max() <= TOL
    cm2 = CoordMap(str(tmp_path) + "/", 1.0, 2, 5)
    assert np.abs(cm2.coords - gold["load_matrix_rot_2_5"]).max() <= 1e-12
    # matrix <-> (xyz, quaternion) helpers round trip
    T = gold["matrices"][3, 4]
    v = H.matrix2xyzquant_torch(T)
    assert np.abs(v.numpy() - gold["load_matrix_rot"][3, 4]).max() <= 1e-12
    assert np.abs(H.xyzquant2matrix_torch(v).numpy() - T).max() <= 1e-6          # float32 container upstream
