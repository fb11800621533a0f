"""In-flight task index: the CPU restatement (oracle/port.cc) against RunningTaskKeeper's
loops run literally over the verbatim TaskDispatcher/RunningTaskBookkeeper (oracle/_ref)."""
import pytest

from running_index_cases import reference_test_case, run_suite


@pytest.mark.parametrize("seed", range(4))
def test_running_index_port_equals_reference(make_dispatcher, seed):
    a = run_suite(make_dispatcher("ref"), seed)
    b = run_suite(make_dispatcher("port"), seed)
    assert len(a) == len(b)
    for k, (x, y) in enumerate(zip(a, b)):
        assert x.shape == y.shape and (x == y).all(), k


@pytest.mark.parametrize("backend", ["port", "ref"])
def test_running_task_keeper_reference_test(make_dispatcher, backend):
    first, second = reference_test_case(make_dispatcher(backend))
    assert first["found"].all() and list(first["servant_task_id"]) == [0, 1, 2]
    assert not second["found"].any()
