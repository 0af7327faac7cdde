"""GPU ROI front-end (td3d_roi_crop_resize) against OpenCV: crop -> cv.resize -> BGR2RGB -> Normalize ->
CHW, as torchdet3d/utils/ie_wrappers.py:155-158,18-21 + utils/transforms.py:16-17 do per box on the host.  Bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import roi_port as rp, torch_port as tp                   # noqa: E402
from test_gpu_model import make_model, t2n, rel, DEV                  # noqa: E402
from torchdet3d_b200 import InferSession                              # noqa: E402
from torchdet3d_b200.preprocess import crop_resize_normalize         # noqa: E402


def _boxes(rng, n, fh, fw, n_frames):
    out = []
    for _ in range(n):
        w, h = int(rng.integers(1, fw)), int(rng.integers(1, fh))
        x0, y0 = int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1))
        out.append((int(rng.integers(0, n_frames)), x0, y0, x0 + w, y0 + h))
    return out


@pytest.mark.parametrize("size", [(224, 224), (320, 320), (96, 128)])
def test_roi_front_end_bit_exact_with_opencv(size):
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    boxes = _boxes(rng, 40, 270, 480, 3) + [(0, 0, 0, 480, 270), (1, 479, 269, 480, 270), (2, 5, 5, 6, 200), (0, 100, 10, 400, 11)]
    got = crop_resize_normalize(torch.tensor(frames, device=DEV), torch.tensor(boxes, dtype=torch.int32, device=DEV), size=size)
    torch.cuda.synchronize()
    for n, (f, x0, y0, x1, y1) in enumerate(boxes):
        ref = rp.reference(frames[f], [(x0, y0, x1, y1)], *size)[0]
        assert np.array_equal(got[n].cpu().numpy(), ref), (n, boxes[n])
    # single-frame form, RGB input (no swap), other constants
    g2 = crop_resize_normalize(torch.tensor(frames[1], device=DEV), torch.tensor([b[1:] for b in boxes[:5]], dtype=torch.int32, device=DEV),
                               size=size, mean=(0.5, 0.4, 0.3), std=(0.2, 0.3, 0.4), bgr=False)
    for n, b in enumerate(boxes[:5]):
        ref = rp.reference(frames[1], [b[1:]], *size, mean=(0.5, 0.4, 0.3), std=(0.2, 0.3, 0.4), bgr=False)[0]
        assert np.array_equal(g2[n].cpu().numpy(), ref), n
    # an empty box is a defined all-zero-pixel crop
    e = crop_resize_normalize(torch.tensor(frames, device=DEV), torch.tensor([(0, 10, 10, 10, 50)], dtype=torch.int32, device=DEV), size=size)
    m, inv = rp.norm_constants()
    assert np.allclose(e[0, :, 0, 0].cpu().numpy(), (0 - m) * inv)


def test_rois_to_keypoints_through_the_session():
    """frames + boxes -> InferSession.load_rois -> keypoints, against oracle crops -> oracle export forward."""
    name = "mobilenetv3_small"
    case = dict(model=name, optim=dict(name="adam"), loss=None)
    _, model = make_model(case)
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (2, 200, 300, 3), dtype=np.uint8)
    boxes = _boxes(rng, 12, 200, 300, 2)
    sess = InferSession(model, 12, 96, 96, chunk=8)
    fr, bx = torch.tensor(frames, device=DEV), torch.tensor(boxes, dtype=torch.int32, device=DEV)
    for _ in range(3):
        sess.load_rois(fr, bx)
        kp, labels, logits = sess.run()
    crops = np.stack([rp.reference(frames[b[0]], [b[1:]], 96, 96)[0] for b in boxes])
    assert np.array_equal(sess.imgs.cpu().numpy(), crops)
    kp_all, lg = tp.forward_export(tp.synth_state(name, seed=0), name, torch.tensor(crops))
    ref_sel, ref_lab = tp.select_by_argmax(kp_all, lg)
    assert np.array_equal(labels.cpu().numpy(), ref_lab.numpy()) and rel(t2n(kp), ref_sel.numpy()) < 1e-3
