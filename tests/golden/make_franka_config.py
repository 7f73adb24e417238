"""Generates tests/golden/franka_config.json from the reference's franka_config.xml (run in the build container, where
/root/reference exists): per-dof position / velocity bounds and observation-noise amplitudes that FrankaRobot reads at
construction (envs/franka_kitchen/franka_env.py:172-202, utils.py).  The oracle Kitchen env loads the JSON, so its tests do not
need the reference checkout."""
import json
import xml.etree.ElementTree as ET

SRC = "/root/reference/gymnasium_robotics/envs/assets/kitchen_franka/franka_assets/franka_config.xml"
root = ET.parse(SRC).getroot()
out = {"name": root.get("name"), "pos_bound": [], "vel_bound": [], "pos_noise_amp": [], "vel_noise_amp": []}
i = 0
while root.find(f"qpos{i}") is not None:
    n = root.find(f"qpos{i}")
    out["pos_bound"].append([float(x) for x in n.get("pos_bound").split()])
    out["vel_bound"].append([float(x) for x in n.get("vel_bound").split()])
    out["pos_noise_amp"].append(float(n.get("pos_noise_amp").split()[0]))
    out["vel_noise_amp"].append(float(n.get("vel_noise_amp").split()[0]))
    i += 1
json.dump(out, open(__file__.replace("make_franka_config.py", "franka_config.json"), "w"), indent=0)
print(i, "dofs")
