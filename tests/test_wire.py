"""FlareStd wire front end on the CPU backends: frames encoded/decoded by the google.protobuf
runtime against the hand-written codec, handlers checked against direct calls, and the verbatim
reference dispatcher (`ref`) against the restatement (`port`) frame by frame."""
import pytest

pytest.importorskip("google.protobuf")

from wire_cases import run_wire_scenario


def test_wire_scenario_port_equals_reference(make_dispatcher):
    a = run_wire_scenario(make_dispatcher, "ref")
    b = run_wire_scenario(make_dispatcher, "port")
    assert a == b
