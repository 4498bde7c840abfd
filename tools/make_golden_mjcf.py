"""Golden constants of the model compiler: parse the REFERENCE's own robot files
(/root/reference/safe_rl_envs/safe_rl_envs/xmls/*.xml, read only here) with
guardx_b200.mjcf.parse_mjcf, append the goal/hazard bodies as World.build does and compile.
tests/test_art_cpu.py checks the built-in descriptions (guardx_b200/robots.py, which need no
XML at run time) against these numbers on machines where /root/reference does not exist.

    python tools/make_golden_mjcf.py        # writes tests/golden/mjcf_constants.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guardx_b200 import mjcf, robots  # noqa: E402

REF = "/root/reference/safe_rl_envs/safe_rl_envs"
FIELDS = ("body_parentid body_pos body_quat body_ipos body_mass body_inertia_full body_invweight0 jnt_type jnt_bodyid "
          "jnt_pos jnt_axis jnt_limited jnt_range jnt_stiffness qpos_spring dof_damping dof_armature dof_invweight0 "
          "geom_type geom_bodyid geom_pos geom_quat geom_size geom_friction geom_margin geom_contype geom_conaffinity "
          "actuator_trnid actuator_gear actuator_ctrllimited actuator_ctrlrange actuator_forcelimited actuator_forcerange "
          "actuator_gainprm0 actuator_biasprm qpos0 body_weldid").split()


def constants(m):
    out = {k: np.asarray(getattr(m, k)).tolist() for k in FIELDS}
    out.update(nq=m.nq, nv=m.nv, nu=m.nu, nbody=m.nbody, timestep=m.opt['timestep'], meaninertia=m.stat_meaninertia,
               contact_pairs=[list(map(int, p)) for p in m.contact_pairs], jnt_name=m.jnt_name)
    return out


if __name__ == "__main__":
    gold = {}
    for f, key in robots.ROBOT_FILES.items():
        spec = mjcf.parse_mjcf(open(os.path.join(REF, f)).read())
        gold[key] = constants(mjcf.compile_spec(mjcf.add_world_geoms(spec, 0.5, 0.3, 8)))
    with open(os.path.join(ROOT, "tests", "golden", "mjcf_constants.json"), "w") as fh:
        json.dump(gold, fh)
    print({k: (v['nq'], v['nu'], round(v['meaninertia'], 6)) for k, v in gold.items()})
