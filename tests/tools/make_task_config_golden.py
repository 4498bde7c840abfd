"""Writes tests/golden/task_configs.json: what the reference's ``configuration_list(task)``
(safe_rl_libX/guard_utils/safe_rl_env_config.py:6-2568) returns for the five tasks the JAX Engine can run,
by executing that function's own source (lifted out of the file with ``ast``: the module's top-level imports --
gym, torch -- are not needed by it).      python tests/tools/make_task_config_golden.py [/root/reference]"""
import ast
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(REF, "safe_rl_libX", "guard_utils", "safe_rl_env_config.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "configuration_list")
ns = {}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "safe_rl_env_config.py", "exec"), ns)
out = {t: ns["configuration_list"](t) for t in ("Goal_Point_8Hazards", "Goal_Swimmer_8Hazards", "Goal_Ant_8Hazards",
                                                "Goal_Walker_8Hazards", "Goal_Doggo_8Hazards")}
path = os.path.join(ROOT, "tests", "golden", "task_configs.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", path)
