"""Writes tests/golden/maze_known_answers.json: the two numeric known-answer vectors the reference's own test-suite
holds for the hot path's reset/goal logic (tests/envs/maze/test_point_maze.py:20-45 of the reference), copied as data
(inputs + expected outputs), not code.  Run in the build container; the JSON is what travels."""
import json
import os

cases = [
    {"name": "test_reset_cell", "source": "tests/envs/maze/test_point_maze.py:20-31",
     "maze_map": [[1, 1, 1, 1], [1, "r", "r", 1], [1, "r", "g", 1], [1, 1, 1, 1]], "seed": 42,
     "options": {"reset_cell": [1, 2]}, "expect": {"reset_pos": [0.67929896, 0.59868401]}, "decimal": 4},
    {"name": "test_goal_cell", "source": "tests/envs/maze/test_point_maze.py:34-45",
     "maze_map": [[1, 1, 1, 1], [1, "r", "g", 1], [1, "g", "g", 1], [1, 1, 1, 1]], "seed": 42,
     "options": {"goal_cell": [2, 1]}, "expect": {"goal": [-0.36302198, -0.53056078]}, "decimal": 4},
]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "maze_known_answers.json"), "w") as f:
    json.dump(cases, f, indent=1)
