"""Generate tests/golden/overlay.npz from the UNMODIFIED reference overlay code.

Run in the authoring container only (needs /root/reference and cv2):

    python tests/golden/gen/make_overlay_goldens.py

Pinned: `blend_images` (pixtrack/visualization/run_vis_on_poses.py:215-219), `add_pose_axes` -> `draw_axes` ->
`project_3d_to_2d` (:66-112) on small seeded images: the blended image, the six projected axis end points (recorded by
wrapping cv2.line) and the image with the axes drawn by OpenCV.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault('PROJECT_ROOT', '/root/reference')
import make_model3d_goldens  # noqa: E402,F401  (installs the stand-ins for the packages that are absent here)

import cv2  # noqa: E402
from pixtrack.visualization import run_vis_on_poses as vis  # noqa: E402

if __name__ == '__main__':
    rng = np.random.default_rng(0)
    H, W = 96, 128
    query = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    nerf = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    blend = vis.blend_images(query, nerf)
    camera = types.SimpleNamespace(size=np.array([W, H], np.float32), f=np.array([150.0, 150.0], np.float32))
    # camera-in-world pose looking at the axes centre from 0.4 units away, slightly rotated
    a, b = 0.3, -0.2
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    centre = np.array([0.1179, 1.1538, 1.3870])
    pose = np.eye(4)
    pose[:3, :3] = Ry @ Rx
    pose[:3, 3] = centre - pose[:3, :3] @ np.array([0.0, 0.0, 0.12])
    recorded = []
    real_line = cv2.line

    def line(img, p0, p1, color, t):
        recorded.append((np.array(p0), np.array(p1), color, t))
        return real_line(img, p0, p1, color, t)
    vis.cv2.line = line
    with_axes = vis.add_pose_axes(blend.copy(), camera, pose, centre.tolist() + [0])
    vis.cv2.line = real_line
    pts = np.array([[r[0], r[1]] for r in recorded]).reshape(6, 2)
    out = os.path.join(HERE, '..', 'overlay.npz')
    np.savez_compressed(out, query=query, nerf=nerf, blend=blend, pose=pose, centre=centre, cam_size=camera.size,
                        cam_f=camera.f, axes_px=pts.astype(np.int16), colors=np.array([r[2] for r in recorded]),
                        thickness=np.array([r[3] for r in recorded]), with_axes=with_axes)
    print('wrote', os.path.abspath(out), os.path.getsize(out), 'bytes; axes end points', pts.tolist())
