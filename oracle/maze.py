"""numpy restatement of the reference's maze grid math and reset draw order -- TEST INFRASTRUCTURE ONLY.

Follows gymnasium_robotics/envs/maze/maze_v4.py: cell_rowcol_to_xy :135-140, cell_xy_to_rowcol :142-146, make_maze
:148-242 (goal/reset cell collection), MazeEnv.reset :299-358, add_xy_position_noise :360-379, compute_reward
:381-388, compute_terminated :390-398.  Pinned against the reference's own known-answer tests
(tests/envs/maze/test_point_maze.py:20-45), see tests/golden/maze_known_answers.json.
"""
from __future__ import annotations

import math

import numpy as np

RESET, GOAL, COMBINED = "r", "g", "c"


class Maze:
    def __init__(self, maze_map, maze_size_scaling=1.0):
        self.maze_map = maze_map
        self.maze_size_scaling = maze_size_scaling
        self.map_length, self.map_width = len(maze_map), len(maze_map[0])
        self.x_map_center = self.map_width / 2 * maze_size_scaling
        self.y_map_center = self.map_length / 2 * maze_size_scaling
        self.unique_goal_locations, self.unique_reset_locations, self.combined_locations, self.wall_cells = [], [], [], []
        empty = []
        for i in range(self.map_length):
            for j in range(self.map_width):
                cell = maze_map[i][j]
                xy = self.cell_rowcol_to_xy(np.array([i, j]))
                if cell == 1:
                    self.wall_cells.append((i, j))
                elif cell == RESET:
                    self.unique_reset_locations.append(xy)
                elif cell == GOAL:
                    self.unique_goal_locations.append(xy)
                elif cell == COMBINED:
                    self.combined_locations.append(xy)
                elif cell == 0:
                    empty.append(xy)
        if not self.unique_goal_locations and not self.unique_reset_locations and not self.combined_locations:
            self.combined_locations = empty
        elif not self.unique_reset_locations and not self.combined_locations:
            self.unique_reset_locations = empty
        elif not self.unique_goal_locations and not self.combined_locations:
            self.unique_goal_locations = empty
        self.unique_goal_locations += self.combined_locations
        self.unique_reset_locations += self.combined_locations

    def cell_rowcol_to_xy(self, rowcol):
        x = (rowcol[1] + 0.5) * self.maze_size_scaling - self.x_map_center
        y = self.y_map_center - (rowcol[0] + 0.5) * self.maze_size_scaling
        return np.array([x, y])

    def cell_xy_to_rowcol(self, xy):
        i = math.floor((self.y_map_center - xy[1]) / self.maze_size_scaling)
        j = math.floor((xy[0] + self.x_map_center) / self.maze_size_scaling)
        return np.array([i, j])


class MazeResetLogic:
    """MazeEnv.reset (maze_v4.py:299-358) without the simulator: returns (goal, reset_pos)."""

    def __init__(self, maze_map, maze_size_scaling=1.0, position_noise_range=0.25):
        self.maze = Maze(maze_map, maze_size_scaling)
        self.position_noise_range = position_noise_range
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))

    def add_xy_position_noise(self, xy):
        nx = self.np_random.uniform(low=-self.position_noise_range, high=self.position_noise_range) * self.maze.maze_size_scaling
        ny = self.np_random.uniform(low=-self.position_noise_range, high=self.position_noise_range) * self.maze.maze_size_scaling
        return np.array([xy[0] + nx, xy[1] + ny])

    def generate_target_goal(self):
        idx = self.np_random.integers(low=0, high=len(self.maze.unique_goal_locations))
        return self.maze.unique_goal_locations[idx].copy()

    def generate_reset_pos(self, goal):
        reset_pos = goal.copy()
        while np.linalg.norm(reset_pos - goal) <= 0.5 * self.maze.maze_size_scaling:
            idx = self.np_random.integers(low=0, high=len(self.maze.unique_reset_locations))
            reset_pos = self.maze.unique_reset_locations[idx].copy()
        return reset_pos

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        options = options or {}
        if options.get("goal_cell") is not None:
            goal = self.maze.cell_rowcol_to_xy(np.asarray(options["goal_cell"]))
        else:
            goal = self.generate_target_goal()
        goal = self.add_xy_position_noise(goal)
        self._raw_goal = goal
        if options.get("reset_cell") is not None:
            reset_pos = self.maze.cell_rowcol_to_xy(np.asarray(options["reset_cell"]))
        else:
            reset_pos = self.generate_reset_pos(goal)
        reset_pos = self.add_xy_position_noise(reset_pos)
        return goal, reset_pos


def compute_reward(achieved_goal, desired_goal, reward_type="sparse"):
    d = np.linalg.norm(achieved_goal - desired_goal, axis=-1)
    return np.exp(-d) if reward_type == "dense" else (d <= 0.45).astype(np.float64)


def compute_terminated(achieved_goal, desired_goal, continuing_task=True):
    if not continuing_task:
        return bool(np.linalg.norm(achieved_goal - desired_goal) <= 0.45)
    return False
