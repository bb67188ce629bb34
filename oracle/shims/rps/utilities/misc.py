"""rps.utilities.misc stand-in (oracle only).  SURVEY.md App. A.5."""
import numpy as np
import matplotlib.pyplot as plt  # noqa: F401  (re-exported: scenarios' visualize.py use `plt` via star import)


def generate_initial_conditions(N, spacing=0.3, width=3, height=1.8):
    x_range = int(np.floor(width / spacing))
    y_range = int(np.floor(height / spacing))
    assert x_range != 0 and y_range != 0
    assert x_range * y_range > N, "Cannot fit %d robots on a %dx%d grid" % (N, x_range, y_range)
    choices = np.random.choice(x_range * y_range, N, replace=False)
    poses = np.zeros((3, N))
    for i, c in enumerate(choices):
        x, y = divmod(c, y_range)
        poses[0, i] = x * spacing - width / 2
        poses[1, i] = y * spacing - height / 2
        poses[2, i] = np.random.rand() * 2 * np.pi - np.pi
    return poses


def determine_marker_size(robotarium_instance, marker_size_meters):
    return 1.0


def determine_font_size(robotarium_instance, font_height_meters):
    return 1.0


def at_pose(states, poses, position_error=0.05, rotation_error=0.2):
    pes = np.linalg.norm(states[:2, :] - poses[:2, :], 2, 0)
    res = np.abs(np.arctan2(np.sin(states[2, :] - poses[2, :]), np.cos(states[2, :] - poses[2, :])))
    return np.nonzero((pes <= position_error) & (res <= rotation_error))
