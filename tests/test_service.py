"""SchedulerServiceImpl restatement over the C ABI, on the CPU backends.  The handlers need
flare's RPC controller and protobuf, so scheduler_service_impl.cc is not compilable here: the
rules are pinned on the reference's own test (scheduler_service_impl_test.cc) and on the
handler source line by line; the dispatcher underneath is the verbatim reference (`ref`) or the
restatement (`port`), and both must agree on everything observable."""
import numpy as np
import pytest

import service_cases as SC

CASES = [SC.token_case, SC.token_with_intersection_case, SC.token_without_intersection_case, SC.heartbeat_rules_case,
         SC.lease_flow_case, SC.token_rollout_case]


def _same(a, b):
    if isinstance(a, np.ndarray):
        return a.shape == b.shape and a.dtype == b.dtype and (a == b).all()
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_service_case(make_dispatcher, case):
    a = case(make_dispatcher("ref"))
    b = case(make_dispatcher("port"))
    assert _same(a, b)


def test_service_needs_both_token_lists(make_dispatcher):
    from yadcc_b200.service import SchedulerService

    d = make_dispatcher("port")
    with pytest.raises(ValueError):
        SchedulerService(d, acceptable_user_tokens="", acceptable_servant_tokens="x")
    with pytest.raises(ValueError):
        SchedulerService(d, acceptable_user_tokens="x", acceptable_servant_tokens="")
    # "a," keeps the empty entry: the empty token is then acceptable (token_verifier.cc:61, keep_empty)
    svc = SchedulerService(d, acceptable_user_tokens="a,", acceptable_servant_tokens="s")
    assert svc.get_config("")[0] == 0 and svc.get_config("b")[0] == 1003
