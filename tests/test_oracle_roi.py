"""CPU: the restated OpenCV uint8 bilinear rule of oracle/roi_port.py against cv2 itself, bit for bit (the GPU ROI
front-end is then compared with either)."""
import numpy as np
import pytest

from oracle import roi_port as rp

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(100, 80), (300, 500), (37, 51), (224, 224), (1000, 700), (5, 7), (223, 225), (2, 2), (1, 9)])
@pytest.mark.parametrize("out", [(224, 224), (320, 320), (96, 128)])
def test_restated_bilinear_is_bit_exact_with_opencv(shape, out):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    assert np.array_equal(rp.resize_bilinear_u8(img, *out), cv2.resize(img, (out[1], out[0])))


def test_pipeline_restatement_equals_opencv_pipeline():
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    boxes = [(10, 20, 200, 300), (0, 0, 640, 480), (600, 400, 640, 480), (100, 100, 101, 103)]
    a = rp.crop_resize_normalize(frame, boxes)
    b = rp.reference(frame, boxes)
    assert a.dtype == np.float32 and a.shape == (4, 3, 224, 224) and np.array_equal(a, b)
