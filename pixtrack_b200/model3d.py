"""Model3D point selection and reference-view choice as array tables (host side, numpy / scipy.sparse).

The reference walks Python dicts and sets per frame:
  * `Model3D.get_p3did_to_dbids(dbids, point_selection='all', min_track_length=3)`
    (pixloc/pixloc/localization/model3d.py:49-87) + the xyz stack in `refine_pose_using_features`
    (pixloc/pixloc/localization/base_refiner.py:96): the 3D points observed by the reference image(s) whose track has
    at least 3 images;
  * `extract_covisibility` (pixtrack/utils/hloc_utils.py:28-47): for every image, how many observations it shares
    with every other image;
  * `PixLocPoseTrackerR9.update_reference_ids` (pixtrack/pose_trackers/pixloc_tracker_r9.py:120-143): among the current
    reference and the images sharing more than 50 observations with it, the one whose rotation is closest (geodesic
    distance, pixtrack/utils/pose_utils.py:8-13) to the current pose.
`PointTables` precomputes these once: per-image (point ids, xyz) arrays, a sparse covisibility matrix, a stacked
rotation table; the per-frame work is one small vectorised query.  Input plumbing only -- no device code.
"""
from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np
from scipy import sparse


class PointTables:
    def __init__(self, image_point3D_ids: Mapping[int, np.ndarray], image_R: Mapping[int, np.ndarray],
                 point_xyz: Mapping[int, np.ndarray], point_image_ids: Mapping[int, Sequence[int]],
                 min_track_length: int = 3):
        """image_point3D_ids[image_id]: int array, -1 = keypoint without a 3D point (colmap Image.point3D_ids);
        image_R[image_id]: 3x3 world-to-camera rotation (Image.qvec2rotmat()); point_xyz[p]: xyz;
        point_image_ids[p]: the point's track (Point3D.image_ids, one entry per observation)."""
        self.image_ids = np.array(sorted(image_point3D_ids), dtype=np.int64)
        self._row = {int(i): r for r, i in enumerate(self.image_ids)}
        self.point_ids = np.array(sorted(point_xyz), dtype=np.int64)
        col = {int(p): c for c, p in enumerate(self.point_ids)}
        self.xyz = np.stack([np.asarray(point_xyz[int(p)], np.float64) for p in self.point_ids]) if len(col) else np.zeros((0, 3))
        self.track_len = np.array([len(point_image_ids[int(p)]) for p in self.point_ids], dtype=np.int64)
        self.R = np.stack([np.asarray(image_R[int(i)], np.float64) for i in self.image_ids])
        self.min_track_length = int(min_track_length)

        # observations of image i (one per matched keypoint) and tracks of point p (one per observation)
        obs_r, obs_c = [], []
        self._points_of_image: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
        for i in self.image_ids:
            ids = np.asarray(image_point3D_ids[int(i)], dtype=np.int64)
            ids = ids[ids != -1]
            cols = np.array([col[int(p)] for p in ids], dtype=np.int64)
            obs_r.append(np.full(cols.shape, self._row[int(i)], dtype=np.int64))
            obs_c.append(cols)
            # selection of model3d.py:62-66 + :80-85: first occurrence order of the point ids (dict insertion order)
            _, first = np.unique(cols, return_index=True)
            uniq = cols[np.sort(first)]
            keep = uniq[self.track_len[uniq] >= self.min_track_length]
            self._points_of_image[int(i)] = (self.point_ids[keep], self.xyz[keep])
        n_i, n_p = len(self.image_ids), len(self.point_ids)
        M = sparse.coo_matrix((np.ones(sum(len(r) for r in obs_r)), (np.concatenate(obs_r), np.concatenate(obs_c))),
                              shape=(n_i, n_p)).tocsr()
        tr_r, tr_c = [], []
        for c, p in enumerate(self.point_ids):
            tr = [self._row[int(j)] for j in point_image_ids[int(p)] if int(j) in self._row]
            tr_r.append(np.full(len(tr), c, dtype=np.int64))
            tr_c.append(np.array(tr, dtype=np.int64))
        L = sparse.coo_matrix((np.ones(sum(len(r) for r in tr_r)), (np.concatenate(tr_r), np.concatenate(tr_c))),
                              shape=(n_p, n_i)).tocsr()
        covis = (M @ L).tolil()
        covis.setdiag(0)
        self.covis = covis.tocsr()
        self.covis.eliminate_zeros()

    # ---- Model3D.get_p3did_to_dbids + xyz stack, for the single-reference case r9 uses (K = 1) ----------------
    def points_of_image(self, image_id: int) -> Tuple[np.ndarray, np.ndarray]:
        """(point ids [n], xyz float64 [n,3]) of the 3D points seen by `image_id` with a long enough track."""
        return self._points_of_image[int(image_id)]

    @property
    def max_points(self) -> int:
        """Largest per-image point set: the capacity a FrameTracker / DeviceEngine needs."""
        return max((len(v[0]) for v in self._points_of_image.values()), default=0)

    def points_of_images(self, image_ids: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        """Union over several reference images, in the reference's first-seen order."""
        ids = np.concatenate([self._points_of_image[int(i)][0] for i in image_ids]) if len(image_ids) else np.zeros(0, np.int64)
        _, first = np.unique(ids, return_index=True)
        ids = ids[np.sort(first)]
        cols = np.searchsorted(self.point_ids, ids)
        return ids, self.xyz[cols]

    # ---- extract_covisibility ---------------------------------------------------------------------------------
    def covisible(self, image_id: int) -> Dict[int, int]:
        row = self.covis.getrow(self._row[int(image_id)])
        return {int(self.image_ids[j]): int(v) for j, v in zip(row.indices, row.data)}

    # ---- update_reference_ids -----------------------------------------------------------------------------------
    def nearest_reference(self, R_query: np.ndarray, current_ref: int, min_covis: int = 50, K: int = 1) -> List[int]:
        """The K images closest in rotation to `R_query` (3x3 world-to-camera) among the current reference and the
        images sharing more than `min_covis` observations with it."""
        r = self._row[int(current_ref)]
        row = self.covis.getrow(r)
        cand = [r] + [int(j) for j, v in zip(row.indices, row.data) if v > min_covis]
        Rd = np.asarray(R_query, np.float64)[None] @ np.transpose(self.R[cand], (0, 2, 1))
        cos = np.clip((np.trace(Rd, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)
        ang = np.arccos(cos)                                   # |rotation vector| of R_q R_ref^T
        order = np.argsort(ang, kind='stable')
        return [int(self.image_ids[cand[o]]) for o in order[:K]]
