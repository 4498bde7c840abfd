"""Config-3 extension, CPU side: the oracle (oracle/task_oracle.py) against closed-form expectations of the DESIGN.md
section-8 specification, and the host-side config lowering.  (CUDA vs oracle: tests/test_task_gpu.py.)"""
import numpy as np
import pytest

from guardx_b200 import task_engine
from oracle.task_oracle import TaskOracle, SLOT_OBJ


def params(cfg):
    c = task_engine.lower(task_engine.parse(cfg))
    return {k: (list(getattr(c, k)) if hasattr(getattr(c, k), '__len__') else getattr(c, k)) for k, _ in c._fields_}


def fresh(cfg, layout_xy, robot_xy, phases=()):
    """oracle with every env placed on the same hand-made layout"""
    P = params(cfg)
    o = TaskOracle(P)
    S = np.zeros((o.n, o.state_dim))
    S[:, SLOT_OBJ:SLOT_OBJ + 2 * o.nobj] = np.asarray(layout_xy, float).reshape(-1)
    S[:, 0:2] = robot_xy
    S[:, 6] = 1.0                      # lagged heading = identity
    for g, ph in enumerate(phases):
        S[:, SLOT_OBJ + 2 * o.nobj + g] = ph
    o.set_state(S)
    obs, dist_task, dist_box, *_ = o.reset_obs()
    o.S[:, 8], o.S[:, 9] = dist_task, dist_box
    return o, P


def test_lowering_and_observation_layout():
    c = task_engine.lower(task_engine.parse({'task': 'push', 'robot_base': 'xmls/point.xml', 'env_num': 5, 'gremlins_num': 0}))
    assert list(c.count) == [1, 1, 0, 4, 4, 2, 0, 1] and c.task == 2 and c.robot == 0 and c.goal_min_robot_dist == 0.0
    assert task_engine.lower(task_engine.parse({'task': 'goal'})).goal_min_robot_dist == 3.0        # engine.py:571
    o = TaskOracle(params({'task': 'push', 'robot_base': 'xmls/point.xml', 'env_num': 2, 'gremlins_num': 0}))
    names = sorted(o.off, key=lambda k: o.off[k][0])
    assert names == sorted(names) and 'gremlins_lidar' not in names and 'box_lidar' in names and 'buttons_lidar' not in names
    assert o.obs_dim == 2 + 16 + 3 + 2 + 16 + 16 + 16 + 3 + 3 + 16
    with pytest.raises(AssertionError, match='Bad key'):
        task_engine.parse({'ghosts_num': 1})
    assert task_engine.configuration('Button_Car_Level2') == {'task': 'button', 'robot_base': 'xmls/car.xml'}


def test_car_straight_line_terminal_speed_and_turning():
    cfg = dict(task='goal', robot_base='xmls/car.xml', env_num=2, hazards_num=1, vases_num=0, pillars_num=0, gremlins_num=0)
    o, P = fresh(cfg, [[50, 50], [40, 40]], [0, 0])
    for _ in range(600):
        o.step(np.ones((2, 2)))
    F = 2 * P['car_gear'] / P['car_wheel_r']
    assert abs(o.S[0, 3] - F / P['car_damp_lin']) < 1e-3 * F / P['car_damp_lin']        # terminal speed F / c
    assert abs(o.S[0, 1]) < 1e-12 and abs(o.S[0, 2]) < 1e-12 and o.S[0, 0] > 5
    o, P = fresh(cfg, [[50, 50], [40, 40]], [0, 0])
    for _ in range(200):
        o.step(np.tile([[-1.0, 1.0]], (2, 1)))
    assert np.abs(o.S[0, 0:2]).max() < 1e-12 and o.S[0, 5] > 0                          # turns on the spot, counter-clockwise
    N = P['car_half_track'] * 2 * P['car_gear'] / P['car_wheel_r']
    assert abs(o.S[0, 5] - N / P['car_damp_ang']) < 2e-2 * N / P['car_damp_ang']


def test_costs_rewards_and_done_rules():
    cfg = dict(task='button', robot_base='xmls/car.xml', env_num=1, buttons_num=2, hazards_num=1, vases_num=1, pillars_num=1,
               gremlins_num=1, num_steps=5)
    # buttons at (1,0) [goal button] and (0,0.25); hazard at (0.1,0); vase at (0,-0.25); pillar at (-0.35,0); gremlin centre (0.3,0.2)
    o, P = fresh(cfg, [[1, 0], [0, 0.25], [0.1, 0], [0, -0.25], [-0.35, 0], [0.3, 0.2]], [0, 0], phases=[np.pi])
    o.S[:, 14] = 0
    obs, r, d, info = o.step(np.zeros((1, 2)))
    rr = P['size'][7]
    hazard = P['size'][3] - 0.1
    grem = np.hypot(0.3 - P['gremlins_travel'], 0.2)           # phase pi: centre - travel along x
    expect = (P['buttons_cost']                                 # wrong button within rr + 0.1
              + hazard + P['vases_contact_cost'] + P['pillars_cost']
              + (P['gremlins_contact_cost'] if grem < rr + P['size'][6] else 0)
              + (P['gremlins_dist_cost'] * (P['gremlins_dist_threshold'] - (grem - P['size'][6])) if grem - P['size'][6] < P['gremlins_dist_threshold'] else 0))
    assert abs(info['cost'][0] - expect) < 1e-12 and abs(expect - 4.2) < 1e-9 and info['violations'][0] == 1 and info['contacts'][0] == 4
    assert d[0] == 0 and abs(r[0]) < 1e-12                      # no motion, no progress
    # pressing the goal button ends the episode; reward is the progress towards it
    o, P = fresh(cfg, [[0.35, 0], [5, 5], [6, 6], [7, 7], [8, 8], [9, 9]], [0, 0])
    o.S[:, 3] = 2.0                                             # moving at 2 m/s towards the goal button
    obs, r, d, info = o.step(np.ones((1, 2)))
    assert d[0] == 0
    obs, r, d, info = o.step(np.ones((1, 2)))                   # the snapshot now sees the robot 0.04 m closer: 0.31 >= 0.3 ...
    for _ in range(3):
        if d[0]:
            break
        obs, r, d, info = o.step(np.ones((1, 2)))
    assert d[0] == 1 and info['flags'][0] & 4 and r[0] > 0
    # timeout: steps > num_steps
    o, P = fresh(cfg, [[3, 3], [5, 5], [6, 6], [7, 7], [8, 8], [9, 9]], [0, 0])
    done_at = None
    for t in range(10):
        obs, r, d, info = o.step(np.zeros((1, 2)))
        if d[0]:
            done_at = t
            break
    assert done_at == 6 and info['flags'][0] == 2               # steps 0..5 pass, the 7th call sees steps = 6 > 5


def test_push_box_and_pillar_projection():
    cfg = dict(task='push', robot_base='xmls/point.xml', env_num=1, hazards_num=0, vases_num=0, pillars_num=1, gremlins_num=0)
    # goal far away, box just ahead (+x) of the Point, pillar to the left
    o, P = fresh(cfg, [[3, 0], [0.25, 0], [0, 0.29]], [0, 0])
    rr, bs, ps = P['size'][7], P['size'][1], P['size'][5]
    o.S[:, 3] = 1.0                                             # moving along +x
    o.S[:, 4] = 0.5                                             # ... and towards the pillar
    o.step(np.array([[1.0, 0.0]]))
    box = o.obj(1, 0)[0]
    q = o.S[0, 0:2]
    assert abs(np.hypot(*(box - q)) - (rr + bs)) < 1e-12        # the box sits at contact distance, pushed along the normal
    assert np.hypot(*(q - np.array([0, 0.29]))) >= rr + ps - 1e-12      # the robot was projected out of the pillar
    # reward of the push task: robot-box progress + box-goal progress
    o, P = fresh(cfg, [[3, 0], [1.0, 0], [9, 9]], [0, 0])
    o.S[:, 3] = 1.0
    o.step(np.array([[1.0, 0.0]]))
    obs, r, d, info = o.step(np.array([[1.0, 0.0]]))
    assert r[0] > 0 and d[0] == 0
