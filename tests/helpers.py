"""Shared test helpers: scene -> oracle inputs, robust comparison metrics."""
import numpy as np

from geosplatting_b200 import scenes
from oracle import raster as R


def oracle_camera(cam: scenes.PinholeCamera) -> R.Camera:
    return R.Camera(cam.view_matrix, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height)


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def linf(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) if np.size(a) else 0.0
