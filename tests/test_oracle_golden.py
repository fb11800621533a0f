"""Pins both CPU checkers on the reference's own golden tests (CPU, no GPU)."""
import pytest

from golden_cases import ALL_CASES


@pytest.mark.parametrize("backend", ["port", "ref"])
@pytest.mark.parametrize("case", ALL_CASES, ids=lambda f: f.__name__)
def test_golden(make_dispatcher, backend, case):
    case(make_dispatcher(backend))
