"""CPU oracle for the point-selection / reference-choice host logic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Dict-and-loop restatement of the reference: `Model3D.get_p3did_to_dbids` with point_selection='all'
(pixloc/pixloc/localization/model3d.py:49-87), `extract_covisibility` (pixtrack/utils/hloc_utils.py:28-47),
`update_reference_ids` (pixtrack/pose_trackers/pixloc_tracker_r9.py:120-143) with
`geodesic_distance_for_rotations` (pixtrack/utils/pose_utils.py:8-13; |rotvec| computed here as arccos((tr-1)/2),
the same angle scipy's as_rotvec returns).
"""
from collections import defaultdict

import numpy as np


def p3did_to_dbids(image_point3D_ids, point_image_ids, dbids, min_track_length=3):
    out = defaultdict(set)
    for dbid in dbids:
        p3dids = np.asarray(image_point3D_ids[dbid])
        for p in p3dids[p3dids != -1]:
            out[int(p)].add(dbid)
    return {i: v for i, v in out.items() if len(point_image_ids[i]) >= min_track_length}


def extract_covisibility(image_point3D_ids, point_image_ids):
    covis_all = {}
    for image_id, ids in image_point3D_ids.items():
        ids = np.asarray(ids)
        covis = defaultdict(int)
        for p in ids[ids != -1]:
            for j in point_image_ids[int(p)]:
                if j != image_id:
                    covis[j] += 1
        if len(covis) == 0:
            continue
        covis_all[image_id] = dict(covis)
    return covis_all


def geodesic(R1, R2):
    Rd = R1 @ R2.T
    return float(np.arccos(np.clip((np.trace(Rd) - 1.0) / 2.0, -1.0, 1.0)))


def update_reference_ids(covis_all, image_R, R_query, current_ref, N=50, K=1):
    covis = covis_all.get(current_ref, {})
    covis = {k: covis[k] for k in covis if covis[k] > N}
    gd = {current_ref: geodesic(R_query, image_R[current_ref])}
    for ref in covis:
        gd[ref] = geodesic(R_query, image_R[ref])
    return sorted(gd, key=lambda x: gd[x])[:K]
