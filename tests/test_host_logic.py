"""Host-side logic that needs no GPU: table composition of the preprocessor point ops, structuring
elements, stream sharding and the world_size-2 detection gather on gloo."""
import os
import socket

import cv2
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cv_ops, synth
from cuauv_vision_pipeline_b200 import sharding, transform
from cuauv_vision_pipeline_b200.preprocessor import point_lut


@pytest.mark.parametrize("rb,gb,bb,con,bri", [(25, 0, 0, 1, 0), (0, -30, 255, 1, 0), (0, 0, 0, 1.7, 0),
                                              (0, 0, 0, 0.35, 0), (0, 0, 0, 1, -40), (10, -20, 30, 2.5, 17),
                                              (-255, 255, 0, 5, -255)])
def test_point_lut_equals_reference_chain(rb, gb, bb, con, bri):
    img = synth.gen_random_bgr(32, 48, 5)
    ref = img
    if rb != 0:
        ref = cv_ops.channel_bias(ref, 2, rb)
    if gb != 0:
        ref = cv_ops.channel_bias(ref, 1, gb)
    if bb != 0:
        ref = cv_ops.channel_bias(ref, 0, bb)
    if con != 1:
        ref = cv_ops.contrast(ref, con)
    if bri != 0:
        ref = cv_ops.brightness(ref, bri)
    lut = point_lut(rb, gb, bb, con, bri)
    got = np.stack([lut[c][img[..., c]] for c in range(3)], axis=-1)
    assert np.array_equal(got, ref)


def test_structuring_elements_match_cv2():
    for x in range(1, 103, 2):
        for y in (x, 1, 3, 21):
            assert np.array_equal(transform.elliptic_kernel(x, y), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (x, y)))
    assert np.array_equal(transform.rect_kernel(5, 3), cv2.getStructuringElement(cv2.MORPH_RECT, (5, 3)))
    with pytest.raises(ValueError):
        transform.elliptic_kernel(4)
    with pytest.raises(ValueError):
        transform.rect_kernel(0)


def test_stream_partition():
    for world in (1, 2, 4, 8):
        owned = [sharding.streams_for_rank(8, r, world) for r in range(world)]
        assert sorted(s for o in owned for s in o) == list(range(8))
        assert all(len(o) == 8 // world for o in owned)
    assert sharding.streams_for_rank(3, 1, 2) == [1]
    with pytest.raises(ValueError):
        sharding.streams_for_rank(8, 2, 2)


def test_merge_in_order():
    assert sharding.merge_in_order([{0: "a", 2: "c"}, {1: "b"}]) == ["a", "b", "c"]
    with pytest.raises(ValueError):
        sharding.merge_in_order([{0: 1}, {0: 2}])
    with pytest.raises(ValueError):
        sharding.merge_in_order([{0: 1}, {2: 2}])
    assert sharding.gather_detections({0: "x"}) == ["x"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sharding.frames_for_rank(7, rank, world)
    local = {f: {"frame": f, "rank": rank, "blobs": [(f, f * 2)]} for f in frames}
    merged = sharding.gather_detections(local, dst=0)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_on_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [m["frame"] for m in merged] == list(range(7))
    assert [m["rank"] for m in merged] == [0, 1, 0, 1, 0, 1, 0]


def test_percentile_replay_equals_numpy():
    """color.thresh_color_distance(auto_distance_percentile=...) takes two order statistics from the device and
    replays numpy's interpolation on the host: same value and same dtype as np.percentile for float32 data."""
    from cuauv_vision_pipeline_b200.color import _percentile_from_order_statistics
    rng = np.random.default_rng(0)
    for t in range(120):
        n = int(rng.integers(1, 5000)) if t % 3 else int(rng.integers(100000, 3000000))
        a = (rng.random(n) ** 2 * 1e4).astype(np.float32)
        q = [50, 1, 99, 0, 100, 33.3, float(rng.uniform(0, 100)), np.float64(12.5), 7][t % 9]
        s = np.sort(a)
        got = _percentile_from_order_statistics(n, q, np.float32, lambda k: s[k])
        ref = np.percentile(a, q)
        assert got == ref and type(got) is type(ref), (n, q, got, ref)


def test_rotation_matrix_equals_cv2():
    import cv2
    from cuauv_vision_pipeline_b200.transform import rotation_matrix_2d
    for ang in (0, 10, -33.3, 90, 180, 359, 0.001, 720.5):
        for c in ((320.0, 240.0), (1104.0, 621.0), (0.5, 7.25)):
            assert np.array_equal(rotation_matrix_2d(c, ang, 1), cv2.getRotationMatrix2D(c, ang, 1))


def test_camera_matrix_yaml_reader():
    """The reference's camera files are OpenCV FileStorage YAML (lib/configs/1_camera_matrix_params.yaml, text copied
    here as data): the reader needs neither OpenCV nor a YAML library."""
    from cuauv_vision_pipeline_b200 import transform
    text = """%YAML:1.0
M: !!opencv-matrix
   rows: 3
   cols: 3
   dt: d
   data: [904.66192735,0.0,481.17596262,0.0,902.84000422,404.82437525,0.0,0.0,1.0]

D: !!opencv-matrix
   rows: 1
   cols: 5
   dt: d
   data: [0.48525658,2.02550297,0.03807578,-0.02152142,-3.30299241]
"""
    m, d = transform.load_camera_matrix_params(text)
    assert m.shape == (3, 3) and m[0, 0] == 904.66192735 and m[1, 2] == 404.82437525 and m[2, 2] == 1.0
    assert d.tolist() == [0.48525658, 2.02550297, 0.03807578, -0.02152142, -3.30299241]
    import pytest
    with pytest.raises(ValueError):
        transform.load_camera_matrix_params("%YAML:1.0\nfoo: 1\n")


def test_optimal_new_camera_matrix_equals_cv2():
    """transform.get_optimal_new_camera_matrix restates cv2.getOptimalNewCameraMatrix in float64 (host-side set-up of the
    undistortion maps, include/camera_filters.hpp:6-11): identical matrices for the reference's camera file and a
    barrel-distorted one, alpha 0 / 0.3 / 1, same and different output sizes."""
    import cv2
    from cuauv_vision_pipeline_b200 import transform
    k = np.array([[904.66192735, 0.0, 481.17596262], [0.0, 902.84000422, 404.82437525], [0.0, 0.0, 1.0]])
    for d in (np.array([0.48525658, 2.02550297, 0.03807578, -0.02152142, -3.30299241]), np.array([-0.25, 0.08, 0.001, -0.0005, 0.0]),
              np.array([-0.3, 0.1, 0.0, 0.0])):
        for size in ((964, 724), (640, 480)):
            for alpha in (0.0, 0.3, 1.0):
                ref, _ = cv2.getOptimalNewCameraMatrix(k, d, size, alpha)
                assert np.array_equal(transform.get_optimal_new_camera_matrix(k, d, size, alpha), ref)
        ref, _ = cv2.getOptimalNewCameraMatrix(k, d, (640, 480), 0.3, (800, 600))
        assert np.array_equal(transform.get_optimal_new_camera_matrix(k, d, (640, 480), 0.3, (800, 600)), ref)
