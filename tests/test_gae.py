"""GAE / rewards-to-go: the oracle restatement against the reference's own formula
(PPOBufferX.finish_path + core.discount_cumsum = scipy.signal.lfilter,
safe_rl_libX/ppo/ppo_train.py:61-113, ppo_core.py:30-46), and the CUDA kernel against
the oracle (bit-exact)."""
import ctypes as C

import numpy as np
import pytest
import scipy.signal


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def discount_cumsum(x, discount):  # ppo_core.py:30-46, verbatim formula
    return scipy.signal.lfilter([1], [1, float(-discount)], x[::-1], axis=0)[::-1]


def reference_finish_paths(rew, val, done, gamma, lam):
    """What ppo()'s loop + PPOBufferX.finish_path produce for an [N, T] buffer when paths
    are closed at done steps and at the epoch end, always with last_val = 0
    (ppo_train.py:388-409)."""
    N, T = rew.shape
    adv = np.zeros((N, T), np.float32)
    ret = np.zeros((N, T), np.float32)
    for i in range(N):
        start = 0
        for t in range(T):
            if done[i, t] > 0 or t == T - 1:
                sl = slice(start, t + 1)
                rews = np.append(rew[i, sl], np.float32(0))
                vals = np.append(val[i, sl], np.float32(0))
                deltas = rews[:-1] + gamma * vals[1:] - vals[:-1]
                adv[i, sl] = discount_cumsum(deltas, gamma * lam).astype(np.float32)
                ret[i, sl] = discount_cumsum(rews, gamma)[:-1].astype(np.float32)
                start = t + 1
    return adv, ret


def _case(seed, T, N, p_done):
    rng = np.random.default_rng(seed)
    rew = rng.normal(size=(T, N)).astype(np.float32)
    val = rng.normal(size=(T, N)).astype(np.float32)
    done = (rng.random((T, N)) < p_done).astype(np.float32)
    return rew, val, done


@pytest.mark.parametrize("T,N,p_done", [(200, 7, 0.02), (50, 33, 0.2), (1, 5, 0.5), (64, 3, 0.0)])
def test_oracle_gae_equals_reference_formula(oracle_lib, T, N, p_done):
    rew, val, done = _case(T * 1000 + N, T, N, p_done)
    gamma, lam = 0.99, 0.97
    adv = np.empty((T, N), np.float32)
    ret = np.empty((T, N), np.float32)
    assert oracle_lib.gxo_gae(_p(rew), _p(val), _p(done), None, T, N, C.c_double(gamma), C.c_double(lam),
                              _p(adv), _p(ret)) == 0
    a_ref, r_ref = reference_finish_paths(rew.T.copy(), val.T.copy(), done.T.copy(), gamma, lam)
    np.testing.assert_array_equal(adv, a_ref.T)
    np.testing.assert_array_equal(ret, r_ref.T)


@pytest.mark.gpu
@pytest.mark.parametrize("T,N,p_done,with_last", [(200, 4096, 0.01, False), (37, 1001, 0.1, True), (1000, 64, 0.003, True)])
def test_cuda_gae_bit_exact(oracle_lib, T, N, p_done, with_last):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200.rollout_ops import gae
    rew, val, done = _case(T + N, T, N, p_done)
    last = np.random.default_rng(1).normal(size=N).astype(np.float32) if with_last else None
    adv = np.empty((T, N), np.float32)
    ret = np.empty((T, N), np.float32)
    oracle_lib.gxo_gae(_p(rew), _p(val), _p(done), _p(last), T, N, C.c_double(0.99), C.c_double(0.95), _p(adv), _p(ret))
    g_adv, g_ret = gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(done).cuda(),
                       None if last is None else torch.from_numpy(last).cuda(), 0.99, 0.95)
    assert np.array_equal(g_adv.cpu().numpy().view(np.uint32), adv.view(np.uint32))
    assert np.array_equal(g_ret.cpu().numpy().view(np.uint32), ret.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("n,p", [(1, 1.0), (1000, 0.3), (4096, 0.0), (70001, 0.02), (65536, 1.0)])
def test_done_compaction(n, p):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200.rollout_ops import done_indices
    rng = np.random.default_rng(n)
    done = (rng.random(n) < p).astype(np.float32)
    idx, count = done_indices(torch.from_numpy(done).cuda())
    want = np.where(done == 1)[0]
    assert int(count.item()) == len(want)
    assert np.array_equal(idx[:len(want)].cpu().numpy(), want)


@pytest.mark.gpu
def test_engine_done_indices():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200 import make
    env = make("Goal_Point_8Hazards", env_num=500, seed=1, max_ep_len=5)
    env.reset()
    for t in range(8):
        o, r, d, i = env.step(torch.zeros(500, 2, device='cuda'))
        idx, count = env.done_indices()
        want = torch.nonzero(d > 0).squeeze(-1).to(torch.int32)
        assert int(count) == want.numel() and torch.equal(idx[:want.numel()], want)
        env.reset_done()
    assert int(count) >= 0
