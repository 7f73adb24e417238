"""Compiles the 4x4 PointMaze model used by the reference's known-answer tests (tests/envs/maze/test_point_maze.py:20-45:
border walls, 2x2 free interior) into tests/golden/pointmaze_4x4.b200m.  Needs /root/reference (build container)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gymnasium_robotics_b200.models import compile_maze_model

MAP = [[1, 1, 1, 1], [1, 0, 0, 1], [1, 0, 0, 1], [1, 1, 1, 1]]
blob = compile_maze_model("point", MAP).to_blob()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pointmaze_4x4.b200m"), "wb").write(blob)
print(len(blob))
