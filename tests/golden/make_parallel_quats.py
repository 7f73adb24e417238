"""Generates tests/golden/parallel_quats.json from the reference's pure-numpy rotation utilities
(gymnasium_robotics/utils/rotations.py:394-408 get_parallel_rotations, :140-159 euler2quat), imported standalone
(the package __init__ needs gymnasium, which is absent here).  Run in the build container only."""
import importlib.util
import json
import os

import numpy as np

spec = importlib.util.spec_from_file_location("ref_rotations", "/root/reference/gymnasium_robotics/utils/rotations.py")
rot = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rot)

quats = [rot.euler2quat(r).tolist() for r in rot.get_parallel_rotations()]
rng = np.random.default_rng(0)
qa, qb = rng.normal(size=(16, 4)), rng.normal(size=(16, 4))
qa /= np.linalg.norm(qa, axis=1, keepdims=True)
qb /= np.linalg.norm(qb, axis=1, keepdims=True)
out = {
    "parallel_quats": quats,
    "quat_mul": {"a": qa.tolist(), "b": qb.tolist(), "out": rot.quat_mul(qa, qb).tolist()},
    "quat_conjugate": rot.quat_conjugate(qa.copy()).tolist(),
    "euler2quat": {"e": qa[:, :3].tolist(), "out": rot.euler2quat(qa[:, :3]).tolist()},
    "quat2euler": {"q": qa.tolist(), "out": rot.quat2euler(qa).tolist()},
}
with open(os.path.join(os.path.dirname(__file__), "parallel_quats.json"), "w") as f:
    json.dump(out, f, indent=0)
print(len(quats))
