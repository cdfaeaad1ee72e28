"""
Drop-in for the hot-path part of the reference's core/geometry.py: compute_separating_vector
(reference core/geometry.py:35-53), in the canonical arithmetic the CUDA kernels use
(sqrt(rn(dx*dx) + rn(dy*dy)), no BLAS), so host-side and device-side normals agree bit for bit.
The reference's other helpers in that file are unused by any caller and are not part of the hot path.
"""
import math

import numpy as np


def compute_separating_vector(ego_pos, obstacle_pos):
    """Unit vector from the ego to the obstacle; [1, 0] when they are closer than 1e-10."""
    d0 = float(obstacle_pos[0]) - float(ego_pos[0])
    d1 = float(obstacle_pos[1]) - float(ego_pos[1])
    norm = math.sqrt(d0 * d0 + d1 * d1)
    if norm < 1e-10:
        return np.array([1.0, 0.0])
    return np.array([d0 / norm, d1 / norm])
