"""GPU ROI front-end: uint8 frames + 2D-detector boxes -> the normalised float32 NCHW crops the regressor consumes.

Replaces the per-box host loop of the reference's deployment wrapper (`Regressor.get_detections` ->
`crop(frame, rect)` -> `IEModel._preprocess` = `cv.resize(crop, (w, h)).transpose(2, 0, 1)`,
torchdet3d/utils/ie_wrappers.py:138-158,18-21) plus BGR->RGB (`utils/transforms.py:10-17`) and the test-time
normalisation (`configs/default_config.py:9-10`) with ONE kernel launch for all boxes (td3d_roi_crop_resize,
csrc/k_roi.cu), bit-exact with OpenCV's uint8 INTER_LINEAR.  Only the uint8 frames cross PCIe (a 224x224 fp32 crop is
602 KB, the 1080p frame it came from 6 MB for all of its boxes)."""
import ctypes as C

import torch

from . import _lib as L

MEAN = (0.5931, 0.4690, 0.4229)      # configs/default_config.py:9
STD = (0.2471, 0.2214, 0.2157)       # configs/default_config.py:10


def _consts(mean, std):
    m = (torch.tensor(mean, dtype=torch.float32) * 255.0).tolist()
    inv = (1.0 / (torch.tensor(std, dtype=torch.float32) * 255.0)).tolist()
    return (C.c_float * 3)(*m), (C.c_float * 3)(*inv)


@torch.no_grad()
def crop_resize_normalize(frames, boxes, size=(224, 224), mean=MEAN, std=STD, bgr=True, out=None):
    """frames: cuda uint8 [H,W,3] or [F,H,W,3]; boxes: cuda int32 [N,4] (x0,y0,x1,y1; single frame) or [N,5]
    (frame index first).  -> float32 [N,3,size[0],size[1]] (written into `out` if given)."""
    if not frames.is_cuda:
        raise L.Td3dError("td3d ROI front-end runs on CUDA (sm_100a) only; there is no CPU fallback")
    assert frames.dtype == torch.uint8 and frames.shape[-1] == 3
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    frames = frames.contiguous()
    boxes = boxes.to(device=frames.device, dtype=torch.int32)
    if boxes.shape[1] == 4:
        boxes = torch.cat([torch.zeros(boxes.shape[0], 1, dtype=torch.int32, device=boxes.device), boxes], dim=1)
    boxes = boxes.contiguous()
    n = boxes.shape[0]
    oh, ow = size
    if out is None:
        out = torch.empty(n, 3, oh, ow, device=frames.device)
    assert out.shape == (n, 3, oh, ow) and out.dtype == torch.float32 and out.is_contiguous()
    m, inv = _consts(mean, std)
    L.check(L.lib().td3d_roi_crop_resize(L.ptr(frames), frames.shape[0], frames.shape[1], frames.shape[2], L.ptr(boxes), n,
                                         oh, ow, m, inv, 1 if bgr else 0, L.ptr(out), L.stream()))
    return out
