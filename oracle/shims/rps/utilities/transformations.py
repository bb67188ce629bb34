"""rps.utilities.transformations stand-in (oracle only).  SURVEY.md App. A.7."""
import numpy as np


def create_si_to_uni_mapping(projection_distance=0.05, angular_velocity_limit=np.pi):
    def si_to_uni_dyn(dxi, poses):
        M, N = np.shape(dxi)
        cs = np.cos(poses[2, :])
        ss = np.sin(poses[2, :])
        dxu = np.zeros((2, N))
        dxu[0, :] = (cs * dxi[0, :] + ss * dxi[1, :])
        dxu[1, :] = (1 / projection_distance) * (-ss * dxi[0, :] + cs * dxi[1, :])
        dxu[1, dxu[1, :] > angular_velocity_limit] = angular_velocity_limit
        dxu[1, dxu[1, :] < -angular_velocity_limit] = -angular_velocity_limit
        return dxu

    def uni_to_si_states(poses):
        _, N = np.shape(poses)
        return poses[:2, :] + projection_distance * np.vstack((np.cos(poses[2, :]), np.sin(poses[2, :])))

    return si_to_uni_dyn, uni_to_si_states
