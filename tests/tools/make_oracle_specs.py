"""Writes tests/golden/oracle_robot_specs.json: oracle.mj_compile.read_xml() of the reference's own robot files
(safe_rl_envs/safe_rl_envs/xmls/{point,swimmer,ant,walker,doggo}.xml), i.e. every element with its attributes merged
over the file's <default> block -- derived data, no reference source text.  The legged oracle compiles its models from
this file at run time (the GPU box has no /root/reference).  Also writes tests/golden/result_xml_structure.json: the
structure of the World.build() dump the reference ships (/root/reference/result.xml: body order, joint names / types /
axes / limited, geom names / types / sizes / contype / conaffinity), which pins the world augmentation of BOTH model
compilers (tests/test_oracle_model.py).

    python tests/tools/make_oracle_specs.py [/root/reference]
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mj_compile  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
XMLS = os.path.join(REF, "safe_rl_envs", "safe_rl_envs", "xmls")


def structure_of_result_xml(path):
    root = ET.parse(path).getroot()
    wb = root.find('worldbody')
    bodies = []
    for body in wb.findall('body'):
        bodies.append({
            'name': body.attrib['name'], 'pos': [float(v) for v in body.attrib['pos'].split()],
            'quat': [float(v) for v in body.attrib['quat'].split()],
            'joints': [{k: j.attrib[k] for k in ('type', 'axis', 'name', 'limited') if k in j.attrib} for j in body.findall('joint')],
            'geoms': [{k: g.attrib[k] for k in ('name', 'type', 'size', 'contype', 'conaffinity') if k in g.attrib}
                      for g in body.findall('geom')]})
    return {'option': dict(root.find('option').attrib), 'bodies': bodies,
            'actuators': [dict(tag=a.tag, **{k: a.attrib[k] for k in ('joint', 'gear') if k in a.attrib}) for a in root.find('actuator')],
            'default': {el.tag: dict(el.attrib) for el in root.find('default') if el.tag in ('joint', 'geom', 'motor', 'velocity')}}


if __name__ == "__main__":
    specs = {}
    for name in ("point", "swimmer", "ant", "walker", "doggo"):
        with open(os.path.join(XMLS, name + ".xml")) as f:
            specs[f"xmls/{name}.xml"] = mj_compile.read_xml(f.read())
    out = os.path.join(ROOT, "tests", "golden", "oracle_robot_specs.json")
    with open(out, "w") as f:
        json.dump(specs, f, indent=0, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")
    out = os.path.join(ROOT, "tests", "golden", "result_xml_structure.json")
    with open(out, "w") as f:
        json.dump(structure_of_result_xml(os.path.join(REF, "result.xml")), f, indent=0, sort_keys=True)
    print("wrote", out)
