"""rps.robotarium stand-in (oracle only).  SURVEY.md App. A.3."""
import numpy as np
from rps.robotarium_abc import RobotariumABC


class Robotarium(RobotariumABC):
    def __init__(self, number_of_robots=-1, show_figure=True, sim_in_real_time=True,
                 initial_conditions=np.array([])):
        super().__init__(number_of_robots, show_figure, sim_in_real_time, initial_conditions)
        self.sim_in_real_time = sim_in_real_time
        self._called_step_already = True
        self._checked_poses_already = False
        self._errors = {}
        self._iterations = 0

    def get_poses(self):
        assert not self._checked_poses_already, "Can only call get_poses() once per call of step()."
        self._called_step_already = False
        self._checked_poses_already = True
        return self.poses

    def call_at_scripts_end(self):
        pass

    def step(self):
        assert not self._called_step_already, "Make sure to call get_poses before calling step() again."
        self._called_step_already = True
        self._checked_poses_already = False

        self._errors = self._validate()
        self._iterations += 1

        p = self.poses
        v = self.velocities
        p[0, :] = p[0, :] + self.time_step * np.cos(p[2, :]) * v[0, :]
        p[1, :] = p[1, :] + self.time_step * np.sin(p[2, :]) * v[0, :]
        p[2, :] = p[2, :] + self.time_step * v[1, :]
        p[2, :] = np.arctan2(np.sin(p[2, :]), np.cos(p[2, :]))
