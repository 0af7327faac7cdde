"""Mirror of the geometry helpers the evaluation needs (reference torchdet3d/utils/geometry.py).

`lift_2d` keeps the reference's signature (a list of 9 x 2 numpy arrays in, a list of 9 x 3 arrays out, geometry.py:51-108) but runs
the whole list as one launch of `td3d_lift_2d` on the GPU; `lift_2d_batch` is the tensor form the evaluation uses.  The small
numpy helpers (camera matrices, NDC conversion, projection) are host code in the reference and stay host code here."""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L

__all__ = ["get_default_camera_matrix", "convert_camera_matrix_2_ndc", "convert_2d_to_ndc", "project_3d_points", "lift_2d",
           "lift_2d_batch"]


def get_default_camera_matrix():
    """geometry.py:16-19"""
    return np.array([[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1]], dtype=np.float64)


def convert_camera_matrix_2_ndc(matrix, img_shape=(1, 1)):
    """geometry.py:29-37"""
    ndc = np.array(matrix, dtype=np.float64, copy=True)
    ndc[0, 0] *= 2.0 / img_shape[0]
    ndc[1, 1] *= 2.0 / img_shape[1]
    ndc[0, 2] = -ndc[0, 2] * 2.0 / img_shape[0] + 1.0
    ndc[1, 2] = -ndc[1, 2] * 2.0 / img_shape[1] + 1.0
    return ndc


def convert_2d_to_ndc(points, portrait=False):
    """geometry.py:40-48"""
    points = np.asarray(points)
    out = np.zeros_like(points)
    if portrait:
        out[:, 0] = points[:, 1] * 2 - 1
        out[:, 1] = points[:, 0] * 2 - 1
    else:
        out[:, 0] = points[:, 0] * 2 - 1
        out[:, 1] = 1 - points[:, 1] * 2
    return out


def project_3d_points(points, camera_matrix):
    """geometry.py:22-26"""
    points = np.asarray(points)
    assert len(points.shape) == 2
    proj = np.matmul(camera_matrix, points.T).T
    proj = proj / -proj[:, 2].reshape(-1, 1)
    return proj[:, :-1]


def _cam_arg(camera_matrix):
    if camera_matrix is None:
        return None, None
    ndc = convert_camera_matrix_2_ndc(camera_matrix)
    arr = (C.c_double * 4)(ndc[0, 0], ndc[1, 1], ndc[0, 2], ndc[1, 2])
    return arr, C.cast(arr, C.c_void_p)


def lift_2d_batch(kp, camera_matrix=None, portrait=False):
    """kp: CUDA float tensor [n, 9, 2] (normalised image coordinates) -> CUDA float64 tensor [n, 9, 3]."""
    L.require_b200()
    assert kp.is_cuda and kp.dim() == 3 and kp.shape[1:] == (9, 2), "lift_2d_batch: expected a CUDA tensor [n, 9, 2]"
    kp = kp.detach().float().contiguous()
    n = kp.shape[0]
    out = torch.empty(n, 9, 3, dtype=torch.float64, device=kp.device)
    keep, cam = _cam_arg(camera_matrix)
    with torch.cuda.device(kp.device):
        L.check(L.lib().td3d_lift_2d(L.ptr(kp), n, 1 if portrait else 0, cam, L.ptr(out), L.stream()))
    del keep
    return out


def lift_2d(keypoint_sets, camera_matrix=None, portrait=False):
    """Reference signature (geometry.py:51-53): list of 9 x 2 arrays -> list of 9 x 3 float64 arrays."""
    if len(keypoint_sets) == 0:
        return []
    for kp in keypoint_sets:
        assert len(kp) == 9
    kp = torch.as_tensor(np.stack([np.asarray(k, dtype=np.float32) for k in keypoint_sets]), device="cuda")
    out = lift_2d_batch(kp, camera_matrix, portrait).cpu().numpy()
    return [out[i] for i in range(out.shape[0])]
