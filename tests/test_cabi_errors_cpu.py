"""CPU: the error / empty-input contract of the C ABI (include/i2p_b200.h), exercised
without a GPU: every call below returns from its argument checks before any kernel launch or CUDA runtime call, so
null pointers are never dereferenced.  The reference has `TORCH_CHECK` in two wrappers and `exit(-1)` on a failed
launch; here every entry point returns a code and leaves a message in i2p_last_error()."""
import pytest

OK, INVALID, UNSUPPORTED = 0, 1, 2
N = None   # a null pointer


@pytest.fixture(scope="module")
def L():
    from i2pnet_b200 import _cabi
    return _cabi.lib()


def test_empty_inputs_are_no_ops(L):
    assert L.i2p_furthest_point_sampling(0, 16, 4, N, N, N, N) == OK            # no clouds
    assert L.i2p_furthest_point_sampling(2, 16, 0, N, N, N, N) == OK            # nothing to sample (sampling_gpu.cu:101)
    assert L.i2p_ball_query(0, 10, 5, 1.0, 4, N, N, N, N) == OK
    assert L.i2p_ball_query(2, 10, 0, 1.0, 4, N, N, N, N) == OK                 # no queries
    assert L.i2p_ball_query(2, 0, 5, 1.0, 4, N, N, N, N) == OK                  # empty cloud: idx keeps the caller's zeros
    assert L.i2p_three_nn(2, 0, 5, N, N, N, N, N) == OK
    assert L.i2p_gather_points(2, 0, 10, 5, N, N, N, N) == OK                   # no channels
    assert L.i2p_group_points(2, 3, 10, 0, 4, N, N, N, N) == OK                 # no centres
    assert L.i2p_group_points(2, 3, 10, 5, 0, N, N, N, N) == OK                 # empty neighbourhoods
    assert L.i2p_gather_rows(2, 100, 8, 0, N, N, N, N) == OK
    assert L.i2p_knn_point(2, 10, 0, 4, N, N, N, N, N) == OK                    # no queries
    assert L.i2p_pw_linear_fwd(0, 8, 16, N, N, N, 1.0, N, N, N, N, N) == OK     # zero rows
    assert L.i2p_quat_mul(0, 5, 1, 5, 0, 0, N, N, N, N) == OK


@pytest.mark.parametrize("call, code, needle", [
    (lambda L: L.i2p_ball_query(2, 10, 5, 1.0, 0, N, N, N, N), INVALID, "ball_query"),                     # nsample < 1
    (lambda L: L.i2p_ball_query(70000, 10, 5, 1.0, 4, N, N, N, N), INVALID, "65535"),                       # grid.y limit
    (lambda L: L.i2p_furthest_point_sampling(1, 0, 4, N, N, N, N), INVALID, "furthest_point_sampling"),    # empty cloud
    (lambda L: L.i2p_furthest_point_sampling(1, 300000, 4, N, N, N, N), UNSUPPORTED, "262144"),            # > cluster capacity
    (lambda L: L.i2p_knn_point(1, 3, 5, 4, N, N, N, N, N), INVALID, "nsample=4 > n=3"),                     # like torch.topk
    (lambda L: L.i2p_knn(1, 5, 100, 64, N, N, N, N, N), UNSUPPORTED, "knn"),                                # k beyond the lane list
    (lambda L: L.i2p_knn(1, 5, 3, 4, N, N, N, N, N), INVALID, "k=4 > m=3"),
    (lambda L: L.i2p_fused_conv_select_k(1, 8, 8, 4, 16, 16, 4, 2, 1.0, 1, 1, N, N, N, N, N, N, N, N, 8, 8, N), INVALID,
     "select"),                                                                                            # 256-cell window > 150 slots
    (lambda L: L.i2p_quat_mul(1, 4, 2, 4, 0, 0, N, N, N, N), INVALID, "quat_mul"),                          # not broadcastable
    (lambda L: L.i2p_pw_linear_fwd(-1, 8, 16, N, N, N, 1.0, N, N, N, N, N), INVALID, "pw_linear_fwd"),
    (lambda L: L.i2p_clip_adam_step(8, N, N, N, N, N, 1e-3, 0.9, 0.999, 1e-8, 0.0, 10.0, 1, N), INVALID, "clip_adam_step"),
])
def test_invalid_arguments_return_codes_and_messages(L, call, code, needle):
    assert call(L) == code
    assert needle in L.i2p_last_error().decode()


def test_host_wrappers_raise_instead_of_exiting():
    """The Python binding turns a non-zero code into I2PError carrying the library's message."""
    from i2pnet_b200 import _cabi
    with pytest.raises(_cabi.I2PError, match="nsample=4 > n=3"):
        _cabi._check(_cabi.lib().i2p_knn_point(1, 3, 5, 4, None, None, None, None, None), "i2p_knn_point")
