"""Regenerates wb_humanoid_mpc_b200/data/g1_wb_model.json and g1_centroidal_model.json from the reference's config files (URDF + task.info + reference.info +
gait.info).  Run in the build container (needs /root/reference); the GPU box only reads the committed JSON."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wb_humanoid_mpc_b200.model_loader import DATA_DIR, build_g1_centroidal_from_reference, build_g1_wb_from_reference, write_flat  # noqa: E402

if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    m = build_g1_wb_from_reference(root)
    DATA_DIR.mkdir(exist_ok=True)
    out = DATA_DIR / "g1_wb_model.json"
    out.write_text(json.dumps(m, indent=1))
    write_flat(m, DATA_DIR / "g1_wb_model.txt")   # the same data for the C++ host layer
    print("wrote", out, "total mass", sum(m["mass"]))
    c = build_g1_centroidal_from_reference(root)
    out = DATA_DIR / "g1_centroidal_model.json"
    out.write_text(json.dumps(c, indent=1))
    write_flat(c, DATA_DIR / "g1_centroidal_model.txt")
    print("wrote", out, "nx", c["nx"], "nu", c["nu"], "torso link on body", c["task_space_cost"]["body"])
