"""Built-in robot descriptions (what ``Robot(robot_base)`` loads in the reference,
world.py:422-445): the kinematic trees, geoms and actuators of the planar-root robots of
``safe_rl_envs/xmls``, written as ``mjcf.Spec`` builders so that no XML file is needed at
run time.  ``tests/test_art_cpu.py`` checks each one against ``mjcf.parse_mjcf`` of the
reference's own file when ``/root/reference`` is present, and against committed golden
constants otherwise.

Robot frame conventions are the reference's: body 'robot' is the root, its first joints
are the planar root (Point/Swimmer: slide-x, slide-y, hinge-z; Ant/Walker: slide-x,
hinge-z, slide-y -- the y slide sits BEHIND the yaw hinge, so it moves along the rotated
body y axis; xmls/ant.xml:19-21, xmls/walker.xml:16-18)."""
from __future__ import annotations

from .mjcf import Spec, OPT_DEFAULT

ROBOT_FILES = {'xmls/point.xml': 'point', 'xmls/ant.xml': 'ant', 'xmls/walker.xml': 'walker',
               'xmls/swimmer.xml': 'swimmer', 'xmls/doggo.xml': 'doggo'}


def _spec(timestep, angle='degree'):
    s = Spec()
    s.opt = dict(OPT_DEFAULT, timestep=timestep)
    s.angle = angle
    s.bodies.append(dict(name='world', parent=-1, pos=(0., 0., 0.), quat=(1., 0., 0., 0.), joints=[], geoms=[]))
    return s


def _body(s, name, parent, pos):
    s.bodies.append(dict(name=name, parent=parent, pos=tuple(float(x) for x in pos), quat=(1., 0., 0., 0.),
                         joints=[], geoms=[]))
    return len(s.bodies) - 1


def point() -> Spec:
    """xmls/point.xml: sphere r=0.1 + box 0.05^3 at x=0.1 (density 1), planar root, three
    'general' actuators (gear 0.3) that inherit a velocity-servo default (assumption A1)."""
    s = _spec(0.02)
    s.defaults['geom'] = dict(condim=3, density=1.0)
    s.defaults['joint'] = dict(damping=0.001)
    s.defaults['actuator'] = dict(ctrlrange=(-1., 1.), ctrllimited=True, forcerange=(-.05, .05), forcelimited=True,
                                  gainprm0=1.0, biasprm=(0., 0., -1.0))
    s.bodies[0]['geoms'].append(dict(name='floor', type='plane', size=(5., 5., 0.1), condim=3))
    r = _body(s, 'robot', 0, (0, 0, 0.1))
    s.bodies[r]['joints'] = [dict(type='slide', axis=(1., 0., 0.), name='robot_x', damping=0.01),
                             dict(type='slide', axis=(0., 1., 0.), name='robot_y', damping=0.01),
                             dict(type='hinge', axis=(0., 0., 1.), name='robot_z', damping=0.005)]
    s.bodies[r]['geoms'] = [dict(name='robot', type='sphere', size=(0.1,), friction=(1., 0.01, 0.01)),
                            dict(name='pointarrow', type='box', pos=(0.1, 0., 0.), size=(0.05, 0.05, 0.05))]
    s.actuators = [dict(joint=f'robot_{a}', gear=0.3, name=a) for a in 'xyz']
    return s


def ant() -> Spec:
    """xmls/ant.xml ('ant_tiny'): planar root at z 0.15, four legs = fixed stub (aux capsule on
    the torso side) + hip (hinge z, +-30 deg) + ankle (horizontal hinge, 30..70 deg magnitude) with
    a foot sphere (the only colliding geoms); 8 motors, gear 70; timestep 0.09."""
    s = _spec(0.09)
    s.defaults['joint'] = dict(armature=1.0, damping=1.0, limited=True)
    s.defaults['geom'] = dict(condim=3, density=5.0, friction=(0.75, 0.01, 0.01), margin=0.01, contype=0, conaffinity=0)
    s.defaults['actuator'] = dict(ctrllimited=True, ctrlrange=(-1., 1.), gear=70.0, gainprm0=1.0, biasprm=(0., 0., 0.))
    s.bodies[0]['geoms'].append(dict(name='floor', type='plane', size=(5., 5., 0.1), conaffinity=1, condim=3))
    r = _body(s, 'robot', 0, (0, 0, 0.15))
    root = dict(armature=0., limited=False, pos=(0., 0., 0.))
    s.bodies[r]['joints'] = [dict(root, type='slide', axis=(1., 0., 0.), damping=0.1, stiffness=0., name='robot_x'),
                             dict(root, type='hinge', axis=(0., 0., 1.), damping=0.01, stiffness=0.1, name='robot_z'),
                             dict(root, type='slide', axis=(0., 1., 0.), damping=0.1, stiffness=0., name='robot_y')]
    s.bodies[r]['geoms'] = [dict(name='robot', type='sphere', pos=(0., 0., 0.), size=(0.06,))]
    legs = [('front_left_leg', 'front_left', 1, 1, (-1., 1., 0.), (30., 70.)),
            ('front_right_leg', 'front_right', -1, 1, (1., 1., 0.), (-70., -30.)),
            ('back_left_leg', 'back_left', -1, -1, (-1., 1., 0.), (-70., -30.)),
            ('back_right_leg', 'back_right', 1, -1, (1., 1., 0.), (30., 70.))]
    for k, (leg, short, sx, sy, ankle_axis, ankle_range) in enumerate(legs, start=1):
        d = (0.05 * sx, 0.05 * sy, 0.0)
        b0 = _body(s, leg, r, (0, 0, 0))
        s.bodies[b0]['geoms'] = [dict(name=f'aux_{k}_geom', type='capsule', fromto=(0., 0., 0., *d), size=(0.02,))]
        b1 = _body(s, f'aux_{k}', b0, d)
        s.bodies[b1]['joints'] = [dict(type='hinge', axis=(0., 0., 1.), name=f'hip_{k}', pos=(0., 0., 0.), range=(-30., 30.))]
        s.bodies[b1]['geoms'] = [dict(name=f'{short}_leg_geom', type='capsule', fromto=(0., 0., 0., *d), size=(0.02,))]
        b2 = _body(s, None, b1, d)
        foot = (0.1 * sx, 0.1 * sy, 0.0)
        s.bodies[b2]['joints'] = [dict(type='hinge', axis=ankle_axis, name=f'ankle_{k}', pos=(0., 0., 0.), range=ankle_range)]
        s.bodies[b2]['geoms'] = [dict(name=f'{short}_ankle_geom', type='capsule', fromto=(0., 0., 0., *foot), size=(0.02,)),
                                 dict(type='sphere', pos=foot, size=(0.02,), contype=1, conaffinity=0)]
    for k in range(1, 5):
        s.actuators += [dict(joint=f'hip_{k}', name=f'hip_{k}'), dict(joint=f'ankle_{k}', name=f'ankle_{k}')]
    return s


def walker() -> Spec:
    """xmls/walker.xml: planar root at z 0.42, two legs of hip (x, z, y hinges), knee, foot;
    foot spheres are the only colliding geoms; 10 motors (gears 10/10/30/30/40); timestep 0.02."""
    s = _spec(0.02)
    s.defaults['joint'] = dict(armature=1.0, damping=20.0, limited=True)
    s.defaults['geom'] = dict(condim=3, density=50.0, friction=(0.75, 0.01, 0.01), contype=0, conaffinity=0)
    s.bodies[0]['geoms'].append(dict(name='floor', type='plane', size=(5., 5., 0.1), conaffinity=1, condim=3))
    r = _body(s, 'robot', 0, (0, 0, 0.42))
    root = dict(armature=0., limited=False, pos=(0., 0., 0.))
    s.bodies[r]['joints'] = [dict(root, type='slide', axis=(1., 0., 0.), damping=10., stiffness=0., name='robot_x'),
                             dict(root, type='hinge', axis=(0., 0., 1.), damping=1., stiffness=0.1, name='robot_z'),
                             dict(root, type='slide', axis=(0., 1., 0.), damping=10., stiffness=0., name='robot_y')]
    s.bodies[r]['geoms'] = [dict(name='robot_geom', type='capsule', fromto=(0., 0.1, 0., 0., -0.1, 0.), size=(0.03,))]
    motor = dict(ctrllimited=True, ctrlrange=(-1., 1.), gainprm0=1.0, biasprm=(0., 0., 0.))
    for side, y, sgn in (('right', 0.1, 1.0), ('left', -0.1, -1.0)):
        hip = dict(armature=0.01, damping=5., stiffness=10., pos=(0., 0., 0.), type='hinge')
        t = _body(s, f'{side}_thigh', r, (0, y, 0))
        s.bodies[t]['joints'] = [dict(hip, axis=(sgn, 0., 0.), name=f'{side}_hip_x', range=(-25., 5.)),
                                 dict(hip, axis=(0., 0., sgn), name=f'{side}_hip_z', range=(-30., 35.)),
                                 dict(hip, axis=(0., 1., 0.), name=f'{side}_hip_y', range=(-100., 10.))]
        s.bodies[t]['geoms'] = [dict(name=f'{side}_thigh_geom', type='capsule', fromto=(0., 0., 0., 0., 0., -0.2), size=(0.03,))]
        g = _body(s, f'{side}_leg', t, (0, 0, -0.3))
        s.bodies[g]['joints'] = [dict(type='hinge', axis=(0., -1., 0.), name=f'{side}_leg_joint', pos=(0., 0., 0.1), range=(-100., 0.))]
        s.bodies[g]['geoms'] = [dict(name=f'{side}_leg_geom', type='capsule', fromto=(0., 0., 0.1, 0., 0., -0.1), size=(0.02,))]
        f = _body(s, f'{side}_foot', g, (0, 0, 0))
        s.bodies[f]['joints'] = [dict(type='hinge', axis=(0., -1., 0.), name=f'{side}_foot_joint', pos=(0., 0., -0.1), range=(-45., 20.))]
        s.bodies[f]['geoms'] = [dict(name=f'{side}_foot_geom', type='capsule', fromto=(0., 0., -0.1, 0.05, 0., -0.1), size=(0.03,), friction=(2.0,)),
                                dict(type='sphere', pos=(0.05, 0., -0.1), size=(0.02,), contype=1, conaffinity=0)]
        s.actuators += [dict(motor, gear=10., joint=f'{side}_hip_x', name=f'{side}_hip_x'),
                        dict(motor, gear=10., joint=f'{side}_hip_z', name=f'{side}_hip_z'),
                        dict(motor, gear=30., joint=f'{side}_hip_y', name=f'{side}_hip_y'),
                        dict(motor, gear=30., joint=f'{side}_leg_joint'),
                        dict(motor, gear=40., joint=f'{side}_foot_joint')]
    return s


def swimmer() -> Spec:
    """xmls/swimmer.xml: three capsule links (density 1000) in the plane z = 0.03, planar root,
    two motors (gear 20) on the link hinges (+-100 deg); timestep 0.03."""
    s = _spec(0.03)
    s.defaults['geom'] = dict(condim=3, density=5.0, friction=(0.75, 0.01, 0.01), margin=0.01, contype=0, conaffinity=0)
    s.defaults['joint'] = dict(armature=0.1)
    s.bodies[0]['geoms'].append(dict(name='floor', type='plane', size=(5., 5., 0.1), conaffinity=1, condim=3))
    r = _body(s, 'robot', 0, (0, 0, 0.03))
    link = dict(type='capsule', density=1000., size=(0.02,), contype=1)
    s.bodies[r]['geoms'] = [dict(link, name='link_1', fromto=(0.3, 0., 0., 0.15, 0., 0.))]
    s.bodies[r]['joints'] = [dict(type='slide', axis=(1., 0., 0.), name='robot_x', pos=(0., 0., 0.)),
                             dict(type='slide', axis=(0., 1., 0.), name='robot_y', pos=(0., 0., 0.)),
                             dict(type='hinge', axis=(0., 0., 1.), name='robot_z', pos=(0., 0., 0.))]
    m = _body(s, 'mid', r, (0.15, 0, 0))
    s.bodies[m]['geoms'] = [dict(link, name='link_2', fromto=(0., 0., 0., -0.15, 0., 0.))]
    s.bodies[m]['joints'] = [dict(type='hinge', axis=(0., 0., 1.), limited=True, name='motor1_rot', pos=(0., 0., 0.), range=(-100., 100.))]
    b = _body(s, 'back', m, (-0.15, 0, 0))
    s.bodies[b]['geoms'] = [dict(link, name='link_3', fromto=(0., 0., 0., -0.15, 0., 0.))]
    s.bodies[b]['joints'] = [dict(type='hinge', axis=(0., 0., 1.), limited=True, name='motor2_rot', pos=(0., 0., 0.), range=(-100., 100.))]
    motor = dict(ctrllimited=True, ctrlrange=(-1., 1.), gear=20.0, gainprm0=1.0, biasprm=(0., 0., 0.))
    s.actuators = [dict(motor, joint='motor1_rot'), dict(motor, joint='motor2_rot')]
    return s


def doggo() -> Spec:
    """xmls/doggo.xml: free-joint torso (two 0.075 capsules; the rear one behind a waist hinge), four
    legs of hip-z, hip-y and ankle hinges with joint springs (stiffness 0.01, springref), every geom a
    colliding capsule; 12 motors, gear 0.02; timestep 0.01.  Runs on the general-tree kernels (gx_tree.cuh)."""
    s = _spec(0.01)
    s.defaults['joint'] = dict(limited=True, stiffness=0.01)
    s.defaults['geom'] = dict(condim=3, size=(0.032,), type='capsule', density=5.0)
    s.defaults['actuator'] = dict(ctrllimited=True, ctrlrange=(-1., 1.), gear=0.02, gainprm0=1.0, biasprm=(0., 0., 0.))
    s.bodies[0]['geoms'].append(dict(name='floor', type='plane', size=(5., 5., 0.1), condim=3))
    r = _body(s, 'robot', 0, (0, 0, 0.3))
    s.bodies[r]['joints'] = [dict(type='free', name='robot')]
    s.bodies[r]['geoms'] = [dict(name='robot', pos=(0., 0., 0.), size=(0.075,), fromto=(0., 0., 0., .2, 0., 0.), type='capsule', density=0.5)]

    def leg(k, parent, sx, sy, hip_y_range, hip_y_spring):
        b0 = _body(s, f'leg_{k}', parent, (0, 0, 0))
        s.bodies[b0]['geoms'] = [dict(name=f'aux_{k}', fromto=(.1 * sx, 0., 0., .2 * sx, .1 * sy, 0.))]
        b1 = _body(s, f'aux_{k}', b0, (.2 * sx, .1 * sy, 0))
        s.bodies[b1]['joints'] = [dict(type='hinge', axis=(0., 0., float(sy)), name=f'hip_{k}_z', pos=(0., 0., 0.), range=(-10., 30.), springref=0.),
                                  dict(type='hinge', axis=(0., 1., 0.), name=f'hip_{k}_y', pos=(0., 0., 0.), range=hip_y_range, springref=hip_y_spring)]
        s.bodies[b1]['geoms'] = [dict(name=f'hip_{k}', fromto=(0., 0., 0., 0.098, 0.0566 * sy, -.05))]
        b2 = _body(s, None, b1, (0.098, 0.0566 * sy, -.05))
        s.bodies[b2]['joints'] = [dict(type='hinge', axis=(-.5 * sy, .866, 0.), name=f'ankle_{k}', pos=(0., 0., 0.), range=(-75., 0.), springref=-20.)]
        s.bodies[b2]['geoms'] = [dict(name=f'ankle_{k}', fromto=(0., 0., 0., -0.1176, -0.0679 * sy, -.1))]

    leg(1, r, 1, 1, (-75., 15.), -10.)
    leg(4, r, 1, -1, (-75., 15.), -10.)
    rear = _body(s, 'rear', r, (0, 0, 0))
    s.bodies[rear]['geoms'] = [dict(name='robot2', pos=(0., 0., 0.), size=(0.075,), fromto=(-.2, 0., 0., 0., 0., 0.), type='capsule', density=0.5)]
    s.bodies[rear]['joints'] = [dict(type='hinge', axis=(1., 0., 0.), name='waist_x', pos=(0., 0., 0.), range=(-30., 30.), springref=0.)]
    leg(2, rear, -1, 1, (0., 135.), 0.)
    leg(3, rear, -1, -1, (0., 135.), 0.)
    for nm in ('hip_{}_z', 'hip_{}_y', 'ankle_{}'):
        s.actuators += [dict(joint=nm.format(k), name=nm.format(k)) for k in (1, 2, 3, 4)]
    return s


BUILTIN = {'point': point, 'ant': ant, 'walker': walker, 'swimmer': swimmer, 'doggo': doggo}


def robot_spec(robot_base: str) -> Spec:
    """``robot_base`` as the reference config spells it ('xmls/ant.xml', engine.py:113)."""
    key = ROBOT_FILES.get(robot_base, robot_base)
    if key not in BUILTIN:
        raise AssertionError(f'robot_base {robot_base!r}: this build has {sorted(ROBOT_FILES)}')
    return BUILTIN[key]()
