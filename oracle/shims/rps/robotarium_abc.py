"""rps.robotarium_abc stand-in (oracle only).  SURVEY.md App. A.1/A.2/A.4."""
import numpy as np
import rps.utilities.misc as misc

# --- constants of RobotariumABC.__init__ (App. A.1); COLLISION_DIAMETER is the flagged one ---
TIME_STEP = 0.033
ROBOT_DIAMETER = 0.11
WHEEL_RADIUS = 0.016
BASE_LENGTH = 0.105
MAX_LINEAR_VELOCITY = 0.2
COLLISION_DIAMETER = 0.135
# Two forms of the collision test exist in published rps revisions and the pinned commit (6bb184e) cannot be read
# here: centre to centre (offset 0: SURVEY.md App. A.4, what every default fixture uses) and the points projected
# 0.025 m along each robot's heading.  ref_harness.RefEnv sets the instance attribute from the config keys
# rps_collision_offset / rps_collision_diameter.
COLLISION_OFFSET = 0.0
BOUNDARIES = [-1.6, -1, 3.2, 2]


class RobotariumABC(object):
    def __init__(self, number_of_robots=-1, show_figure=True, sim_in_real_time=True,
                 initial_conditions=np.array([])):
        assert isinstance(number_of_robots, int)
        assert isinstance(initial_conditions, np.ndarray)
        assert 0 <= number_of_robots <= 50
        if initial_conditions.size > 0:
            assert initial_conditions.shape == (3, number_of_robots)

        self.number_of_robots = number_of_robots
        self.show_figure = show_figure
        self.initial_conditions = initial_conditions
        self.boundaries = list(BOUNDARIES)

        self.time_step = TIME_STEP
        self.robot_diameter = ROBOT_DIAMETER
        self.wheel_radius = WHEEL_RADIUS
        self.base_length = BASE_LENGTH
        self.max_linear_velocity = MAX_LINEAR_VELOCITY
        self.max_angular_velocity = 2 * (self.wheel_radius / self.robot_diameter) * \
            (self.max_linear_velocity / self.wheel_radius)
        self.max_wheel_velocity = self.max_linear_velocity / self.wheel_radius
        self.robot_radius = self.robot_diameter / 2
        self.collision_diameter = COLLISION_DIAMETER
        self.collision_offset = COLLISION_OFFSET

        self.velocities = np.zeros((2, number_of_robots))
        self.poses = self.initial_conditions        # NO copy: callers alias the simulator state
        if self.initial_conditions.size == 0:
            self.poses = misc.generate_initial_conditions(self.number_of_robots, spacing=0.2,
                                                          width=2.5, height=1.5)
        self.figure = None
        self.axes = None

    def set_velocities(self, ids, velocities):
        idxs = np.where(np.abs(velocities[0, :]) > self.max_linear_velocity)
        velocities[0, idxs] = self.max_linear_velocity * np.sign(velocities[0, idxs])
        idxs = np.where(np.abs(velocities[1, :]) > self.max_angular_velocity)
        velocities[1, idxs] = self.max_angular_velocity * np.sign(velocities[1, idxs])
        self.velocities = velocities

    def _uni_to_diff(self, dxu):
        r = self.wheel_radius
        l = self.base_length
        dxdd = np.vstack((1 / (2 * r) * (2 * dxu[0, :] - l * dxu[1, :]),
                          1 / (2 * r) * (2 * dxu[0, :] + l * dxu[1, :])))
        return dxdd

    def _validate(self, errors={}):
        # The mutable default is deliberate: in rps one dict is shared by every call and every
        # Robotarium instance of the process; MARBLER's roboEnv.py:80-91 relies on the counts being
        # cumulative across instances.
        p = self.poses
        b = self.boundaries
        N = self.number_of_robots

        for i in range(N):
            x = p[0, i]
            y = p[1, i]
            if x < b[0] or x > (b[0] + b[2]) or y < b[1] or y > (b[1] + b[3]):
                if "boundary" in errors:
                    if i in errors["boundary"]:
                        errors["boundary"][i] += 1
                    else:
                        errors["boundary"][i] = 1
                else:
                    errors["boundary"] = {i: 1}
                    errors["boundary_string"] = "iteration(s) robots were outside the boundaries."

        for j in range(N - 1):
            for k in range(j + 1, N):
                first_position = p[:2, j] + self.collision_offset * np.array([np.cos(p[2, j]), np.sin(p[2, j])])
                second_position = p[:2, k] + self.collision_offset * np.array([np.cos(p[2, k]), np.sin(p[2, k])])
                if np.linalg.norm(first_position - second_position) <= self.collision_diameter:
                    if "collision" in errors:
                        if j in errors["collision"]:
                            errors["collision"][j] += 1
                        else:
                            errors["collision"][j] = 1
                    else:
                        errors["collision"] = {j: 1}
                        errors["collision_string"] = "iteration(s) where robots collided."

        dxdd = self._uni_to_diff(self.velocities)
        if np.any(np.absolute(dxdd) > self.max_wheel_velocity):
            if "actuator" in errors:
                errors["actuator"] += 1
            else:
                errors["actuator"] = 1
                errors["actuator_string"] = "iteration(s) where the actuator limits were exceeded."
        return errors
